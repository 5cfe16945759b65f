// goicp_device.cuh -- device-side building blocks shared by the Go-ICP kernels (sm_100a).
//
// Numerics contract (see DESIGN.md "Parity"): everything that feeds a voxel index or a
// comparison is computed with the reference's operation order and roundings.  The library
// is compiled with -fmad=false and the index path additionally uses explicit _rn
// intrinsics, because the reference object code (x86-64, no -mfma) contains no fused
// multiply-adds (SURVEY.md section 7 "libm / FMA parity").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace goicp {

// Distance-transform grid resident in HBM/L2: S^3 floats in the reference's [z][y][x]
// order (jly_3ddt.h:27-79) plus the frame of DT3D (jly_3ddt.h:100-111).
struct DtView {
    const float* __restrict__ grid;
    int S;
    int S2;                 // S*S
    double xmin, ymin, zmin, scale;
    double inv_scale;       // 1/scale, used only by the search kernels' tree-sum path (see accumulate_point8)
    cudaTextureObject_t tex; // the same grid as a 1-D linear texture: the bound kernels gather through the TEX pipe (engine.cu: dt_texture); 0 = none, plain loads
};

// One axis of DT3D::Distance (jly_3ddt.cpp:983-1016): idx = int((q - min)*scale + 0.5) in
// double with truncation toward zero; `over` is the signed overshoot in voxels (0 inside).
__device__ __forceinline__ void dt_axis(float q, double mn, double scale, int S, int& idx, float& over)
{
    double v = __dadd_rn(__dmul_rn(__dsub_rn((double)q, mn), scale), 0.5);
    int i = __double2int_rz(v);
    float o = 0.0f;
    if (i < 0) { o = (float)i; i = 0; }
    else if (i >= S) { o = (float)(i - S + 1); i = S - 1; }
    idx = i; over = o;
}
__device__ __forceinline__ int dt_axis_raw(float q, double mn, double scale)
{
    return __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)q, mn), scale), 0.5));
}
// Out-of-grid correction (jly_3ddt.cpp:1025): float sqrt of the float overshoot norm, then
// a double divide and a double add with the clamped voxel's distance, rounded to float.
__device__ __forceinline__ float dt_outside(float d, float a, float b, float c, double scale)
{
    float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
    return __double2float_rn(__dadd_rn(__ddiv_rn((double)e, scale), (double)d));
}
__device__ __forceinline__ float dt_distance(const DtView& dt, float qx, float qy, float qz)
{
    int ix, iy, iz; float a, b, c;
    dt_axis(qx, dt.xmin, dt.scale, dt.S, ix, a);
    dt_axis(qy, dt.ymin, dt.scale, dt.S, iy, b);
    dt_axis(qz, dt.zmin, dt.scale, dt.S, iz, c);
    float d = __ldg(dt.grid + (size_t)iz * dt.S2 + iy * dt.S + ix);
    if (a != 0.0f || b != 0.0f || c != 0.0f) d = dt_outside(d, a, b, c, dt.scale);
    return d;
}

// R*p exactly as jly_goicp.cpp:470-476 / :105-107: ((r0*x + r1*y) + r2*z), every product and
// sum rounded to float.
__device__ __forceinline__ float dot3_ref(float r0, float r1, float r2, float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(r0, x), __fmul_rn(r1, y)), __fmul_rn(r2, z));
}

// ---- warp reduction of 16 per-lane values in 16 shuffles --------------------------------
// After the call, lane L holds in v[0] the warp-wide sum of value index ((L >> 1) & 15).
__device__ __forceinline__ void warp_reduce16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float send = hi ? v[k] : v[k + 8];
            float keep = hi ? v[k + 8] : v[k];
            v[k] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float send = hi ? v[k] : v[k + 4];
            float keep = hi ? v[k + 4] : v[k];
            v[k] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float send = hi ? v[k] : v[k + 2];
            float keep = hi ? v[k + 2] : v[k];
            v[k] = keep + __shfl_xor_sync(full, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        float send = hi ? v[0] : v[1];
        float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 2);
    }
    v[0] = v[0] + __shfl_xor_sync(full, v[0], 1);
}

// ---- one data point against the 8 children of a translation cube -------------------------
// P = rotated point (x,y,z) and its rotation-uncertainty radius gamma (0 in the upper-bound
// pass).  tr[0..1]/[2..3]/[4..5] = x/y/z translation of the children whose axis bit is 0/1
// (jly_goicp.cpp:265-272), tr[6] = maxTransDis (:263); tr may live in shared memory.
// acc[j] += d^2 (ub, :302-306), acc[8+j] += max(d-gt,0)^2 (lb, :308-315) with
// d = max(DT(P+t_j) - gamma, 0) (:276-291).  The three per-axis voxel indices are shared between
// children: 6 double-precision index computations serve 8 gathers.  The in-grid case (all six
// indices inside) is the fast path; anything else goes through the reference's clamp +
// overshoot formula per child.
__device__ __forceinline__ void point_residuals8(const DtView& dt, float px, float py, float pz, float gamma,
                                                 const float* __restrict__ tr, float (&m)[8])
{
    // per axis and axis bit: clamped voxel index and squared overshoot (0 inside the grid);
    // overshoot a = x (x<0) or x-S+1 (x>=S) == raw - clamped (jly_3ddt.cpp:991-1023)
    int xi[2], yo[2], zo[2];
    float ax2[2], ay2[2], az2[2];
    const int Sm1 = dt.S - 1;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        int i = dt_axis_raw(__fadd_rn(px, tr[b]), dt.xmin, dt.scale);
        int cl = min(max(i, 0), Sm1); float a = (float)(i - cl);
        xi[b] = cl; ax2[b] = __fmul_rn(a, a);
        i = dt_axis_raw(__fadd_rn(py, tr[2 + b]), dt.ymin, dt.scale);
        cl = min(max(i, 0), Sm1); a = (float)(i - cl);
        yo[b] = cl * dt.S; ay2[b] = __fmul_rn(a, a);
        i = dt_axis_raw(__fadd_rn(pz, tr[4 + b]), dt.zmin, dt.scale);
        cl = min(max(i, 0), Sm1); a = (float)(i - cl);
        zo[b] = cl * dt.S2; az2[b] = __fmul_rn(a, a);
    }
    float d[8];
#pragma unroll
    if (dt.tex) {
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = tex1Dfetch<float>(dt.tex, zo[(j >> 2) & 1] + yo[(j >> 1) & 1] + xi[j & 1]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = __ldg(dt.grid + (zo[(j >> 2) & 1] + yo[(j >> 1) & 1] + xi[j & 1]));
    }
    if ((ax2[0] + ax2[1] + ay2[0] + ay2[1] + az2[0] + az2[1]) != 0.0f) {
        // some child of this point leaves the grid: sqrt(a^2+b^2+c^2)/scale + clamped distance
        // (jly_3ddt.cpp:1025).  The search path multiplies by 1/scale instead of dividing in double
        // (a 1-ulp(double) difference before the final rounding to float, far below the float
        // summation-order noise of this path); the strict paths divide exactly (dt_outside).
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float s2 = __fadd_rn(__fadd_rn(ax2[j & 1], ay2[(j >> 1) & 1]), az2[(j >> 2) & 1]);
            if (s2 != 0.0f) d[j] = __double2float_rn(__dadd_rn(__dmul_rn((double)__fsqrt_rn(s2), dt.inv_scale), (double)d[j]));
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float v = __fsub_rn(d[j], gamma);
        m[j] = v < 0.0f ? 0.0f : v;
    }
}

__device__ __forceinline__ void accumulate_point8(const DtView& dt, float px, float py, float pz, float gamma,
                                                  const float* __restrict__ tr, float (&acc)[16])
{
    float m[8];
    point_residuals8(dt, px, py, pz, gamma, tr, m);
    const float gt = tr[6];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        acc[j] = __fadd_rn(acc[j], __fmul_rn(m[j], m[j]));
        const float e = __fsub_rn(m[j], gt);
        if (e > 0.0f) acc[8 + j] = __fadd_rn(acc[8 + j], __fmul_rn(e, e));
    }
}

// Same for a single translation (generic pair evaluation / pose scoring).
__device__ __forceinline__ void accumulate_point1(const DtView& dt, float px, float py, float pz, float gamma,
                                                  float tx, float ty, float tz, float gt, float& ub, float& lb)
{
    float d = dt_distance(dt, __fadd_rn(px, tx), __fadd_rn(py, ty), __fadd_rn(pz, tz));
    float m = __fsub_rn(d, gamma);
    m = m < 0.0f ? 0.0f : m;
    ub = __fadd_rn(ub, __fmul_rn(m, m));
    float e = __fsub_rn(m, gt);
    if (e > 0.0f) lb = __fadd_rn(lb, __fmul_rn(e, e));
}

// maxTransDis = float(SQRT3/2.0 * w) in double (jly_goicp.cpp:263 with SQRT3 = 1.732050808)
__device__ __forceinline__ float max_trans_dis(float w)
{
    return __double2float_rn(__dmul_rn(1.732050808 / 2.0, (double)w));
}

} // namespace goicp
