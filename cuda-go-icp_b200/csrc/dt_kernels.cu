// dt_kernels.cu -- distance-transform builders on sm_100a (GoICP::BuildDT, jly_goicp.cpp:75-90).
//
// GOICP_DT_REFERENCE reproduces DT3D::Build (jly_3ddt.cpp:889-979) bit for bit.  The reference
// transform is a *sequential* signed-less vector propagation ("DEuclidean", :710-742): every
// voxel update reads voxels updated earlier in the same raster scan, each x-slice is swept by
// rows forward then backward, each row along z forward then backward, and the whole volume in
// +x then -x.  Its output depends on that order (it is not the exact EDT), so the order is
// kept: one persistent CTA walks slices and rows in the reference order; within a row all the
// candidates that do not depend on the running scan are evaluated in parallel (one thread per
// voxel) and only the true recurrence -- "previous voxel of this scan + (0,0,1)" -- is resolved
// serially, and only for rows where it can matter.  Comparisons use the integer squared norm:
// the reference compares float(sqrt(v^2+h^2+d^2)), a strictly increasing function of that
// integer for grids up to 1024^3, with strict '<' in mask order (first minimum wins).
//
// GOICP_DT_EXACT_EDT is the separable exact squared-Euclidean transform (three 1-D lower-envelope
// passes over integer squared distances), fully parallel.
#include "dt_kernels.h"
#include <cmath>
#include <cstdint>
#include <vector>

namespace goicp {

namespace {

constexpr int kInf = 0x3fffffff;       // "no candidate" squared norm
constexpr int kMaxS = 1024;

struct __align__(8) Vox { short v, h, d, pad; };

__device__ __forceinline__ int vox_norm2(const Vox& a) { return (int)a.v * a.v + (int)a.h * a.h + (int)a.d * a.d; }
__device__ __forceinline__ bool vox_unset(const Vox& a) { return a.v == 32767; }

__global__ void dt_init_kernel(Vox* G, size_t n3)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    Vox u; u.v = u.h = u.d = 32767; u.pad = 0;
    // As compiled (g++ 13.3 -O2) the reference's mask function returns (0,0,0) for the very first
    // voxel it visits, where no candidate qualifies and its result struct is uninitialised
    // (jly_3ddt.cpp:469-470): voxel (0,0,0) acts as one extra seed.  Pinned against oracle/_ref.
    if (i == 0) u.v = u.h = u.d = 0;
    G[i] = u;
}

// seeds: ROUND((p - min)*scale) in double, points outside the grid skipped (jly_3ddt.cpp:952-966)
__global__ void dt_seed_kernel(Vox* G, int S, const float* __restrict__ model, int nm, double xmin, double ymin, double zmin, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    int x = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i], xmin), scale), 0.5));
    int y = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 1], ymin), scale), 0.5));
    int z = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 2], zmin), scale), 0.5));
    if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) return;
    Vox s; s.v = s.h = s.d = 0; s.pad = 0;
    G[((size_t)x * S + y) * S + z] = s;
}

// ---- the sequential propagation ------------------------------------------------------------
struct Cand { int n2; Vox vec; };

__device__ __forceinline__ void consider(Cand& best, const Vox* G, int S, int x, int y, int z, int iv, int ih, int id)
{
    if ((unsigned)x >= (unsigned)S || (unsigned)y >= (unsigned)S || (unsigned)z >= (unsigned)S) return;
    Vox s = G[((size_t)x * S + y) * S + z];
    if (vox_unset(s)) return;                       // scores ~56755 > 32767 in the reference: never chosen
    s.v = (short)(s.v + iv); s.h = (short)(s.h + ih); s.d = (short)(s.d + id);
    int n2 = vox_norm2(s);
    if (n2 < best.n2) { best.n2 = n2; best.vec = s; }
}

// Row-scan kinds (mask functions of jly_3ddt.cpp and where the running-scan entry sits in the
// tie order): F1 = MINforwardDE1 (:502-706), F3 = MINforwardDE3 (:51-131), B1 = MINbackwardDE1
// (:295-500), B3 = MINbackwardDE3 (:171-252), C_UP = MINforwardDE4 (:133-169, scan z ascending),
// C_DN = MINforwardDE2 (:254-293, scan z descending).
enum ScanKind { F1 = 0, F3 = 1, B1 = 2, B3 = 3, C_UP = 4, C_DN = 5 };

struct RowShared {
    int nP[kMaxS]; int nQ[kMaxS];
    short dP[kMaxS]; short dQ[kMaxS];
    Vox vP[kMaxS]; Vox vQ[kMaxS];
    short start[kMaxS]; unsigned char tag[kMaxS];      // 0 = P, 1 = chain, 2 = Q, 3 = none
};

template <int KIND>
__device__ void row_scan(Vox* G, int S, int x, int y, RowShared& sh)
{
    const int z = threadIdx.x;
    const bool active = z < S;
    // scan direction: +1 = z ascending (recurrence reads z-1), -1 = z descending (reads z+1)
    constexpr int dir = (KIND == F1 || KIND == B3 || KIND == C_UP) ? +1 : -1;
    Cand P, Q;
    P.n2 = kInf; Q.n2 = kInf; P.vec.v = P.vec.h = P.vec.d = 32767; P.vec.pad = 0; Q.vec = P.vec;
    if (active) {
        if (KIND == F1 || KIND == B1) {
            const int xs = KIND == F1 ? x - 1 : x + 1;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dz = -1; dz <= 1; dz++)
                    consider(P, G, S, xs, y + dy, z + dz, 1, dy != 0, dz != 0);
        }
        if (KIND == F1 || KIND == B3) {              // previous row y-1, then self; recurrence comes last
            consider(P, G, S, x, y - 1, z - 1, 0, 1, 1);
            consider(P, G, S, x, y - 1, z, 0, 1, 0);
            consider(P, G, S, x, y - 1, z + 1, 0, 1, 1);
            consider(P, G, S, x, y, z, 0, 0, 0);
        } else if (KIND == F3 || KIND == B1) {       // recurrence first, then (z,y+1), self, (z-1,y+1)
            consider(Q, G, S, x, y + 1, z, 0, 1, 0);
            consider(Q, G, S, x, y, z, 0, 0, 0);
            consider(Q, G, S, x, y + 1, z - 1, 0, 1, 1);
        } else {                                     // pure chains: recurrence first, then self
            consider(Q, G, S, x, y, z, 0, 0, 0);
        }
        sh.nP[z] = P.n2; sh.dP[z] = P.vec.d; sh.vP[z] = P.vec;
        sh.nQ[z] = Q.n2; sh.dQ[z] = Q.vec.d; sh.vQ[z] = Q.vec;
    }
    __syncthreads();
    // Static test: would the recurrence win anywhere if its source were the neighbour's own
    // (recurrence-free) result?  If nowhere, it wins nowhere (induction along the scan) and the
    // row is settled without any serial work.
    int win = 0;
    if (active) {
        const int zp = z - dir;
        if (zp >= 0 && zp < S) {
            int n0 = sh.nP[zp], d0 = sh.dP[zp];
            if (sh.nQ[zp] < n0) { n0 = sh.nQ[zp]; d0 = sh.dQ[zp]; }
            if (n0 < kInf) {
                const int nc = n0 + 2 * d0 + 1;
                win = (nc < P.n2) && (nc <= Q.n2);
            }
        }
    }
    const int any = __syncthreads_or(win);
    if (any) {
        if (threadIdx.x == 0) {
            int n_prev = kInf, d_prev = 0, s_prev = 0;
            for (int k = 0; k < S; k++) {
                const int zz = dir > 0 ? k : S - 1 - k;
                const int nP = sh.nP[zz], nQ = sh.nQ[zz];
                const int nc = n_prev < kInf ? n_prev + 2 * d_prev + 1 : kInf;
                int t, n, d, s;
                if (nP <= nc && nP <= nQ) { t = nP < kInf ? 0 : 3; n = nP; d = sh.dP[zz]; s = zz; }
                else if (nc <= nQ)        { t = 1; n = nc; d = d_prev + 1; s = s_prev; }
                else                      { t = 2; n = nQ; d = sh.dQ[zz]; s = zz; }
                sh.tag[zz] = (unsigned char)t; sh.start[zz] = (short)s;
                n_prev = n; d_prev = d; s_prev = s;
            }
        }
        __syncthreads();
    }
    if (active) {
        Vox out;
        bool write = true;
        if (!any) {
            if (P.n2 <= Q.n2) { out = P.vec; write = P.n2 < kInf; } else out = Q.vec;
        } else {
            const int t = sh.tag[z];
            if (t == 0) out = P.vec;
            else if (t == 2) out = Q.vec;
            else if (t == 1) {
                const int s = sh.start[z];
                out = sh.tag[s] == 0 ? sh.vP[s] : sh.vQ[s];
                out.d = (short)(out.d + (dir > 0 ? z - s : s - z));
            } else write = false;
        }
        if (write) G[((size_t)x * S + y) * S + z] = out;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kMaxS)
dt_propagate_kernel(Vox* G, int S)
{
    __shared__ RowShared sh;
    for (int x = 0; x < S; x++) {                                   // jly_3ddt.cpp:719-728
        for (int y = 0; y < S; y++) { row_scan<F1>(G, S, x, y, sh); row_scan<C_DN>(G, S, x, y, sh); }
        for (int y = S - 1; y >= 0; y--) { row_scan<F3>(G, S, x, y, sh); row_scan<C_UP>(G, S, x, y, sh); }
    }
    for (int x = S - 1; x >= 0; x--) {                              // :729-739
        for (int y = S - 1; y >= 0; y--) { row_scan<B1>(G, S, x, y, sh); row_scan<C_UP>(G, S, x, y, sh); }
        for (int y = 0; y < S; y++) { row_scan<B3>(G, S, x, y, sh); row_scan<C_DN>(G, S, x, y, sh); }
    }
}

// distance = float( double(float(sqrt(double(n2)))) / scale ), clamped at 0 (jly_3ddt.cpp:970-978);
// also transposes the working [x][y][z] layout into the reference's [z][y][x].
__global__ void dt_finalize_kernel(const Vox* __restrict__ G, int S, double scale, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    {
        const int x = x0 + threadIdx.y, z = z0 + threadIdx.x;
        if (x < S && z < S) {
            const Vox a = G[((size_t)x * S + y) * S + z];
            const double n2 = (double)a.v * (double)a.v + (double)((int)a.h * a.h) + (double)((int)a.d * a.d);
            float dv = __double2float_rn(sqrt(n2));
            float r = __double2float_rn(__ddiv_rn((double)dv, scale));
            tile[threadIdx.y][threadIdx.x] = r < 0.0f ? 0.0f : r;
        }
    }
    __syncthreads();
    {
        const int x = x0 + threadIdx.x, z = z0 + threadIdx.y;
        if (x < S && z < S) out[((size_t)z * S + y) * S + x] = tile[threadIdx.x][threadIdx.y];
    }
}

// ---- exact EDT (separable, integer squared distances) ----------------------------------------
// pass along z per (x,y) column: squared distance to the nearest seed in the column
__global__ void edt_pass_z(const Vox* __restrict__ G, int S, int* __restrict__ D)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;      // x*S + y
    if (col >= S * S) return;
    const Vox* g = G + (size_t)col * S;
    int* d = D + (size_t)col * S;
    int last = -kMaxS * 4;
    for (int z = 0; z < S; z++) { if (g[z].v == 0 && g[z].h == 0 && g[z].d == 0) last = z; int k = z - last; d[z] = k > 4 * kMaxS - 1 ? kInf : k * k; }
    last = kMaxS * 8;
    for (int z = S - 1; z >= 0; z--) { if (g[z].v == 0 && g[z].h == 0 && g[z].d == 0) last = z; int k = last - z; int v = k > 4 * kMaxS - 1 ? kInf : k * k; if (v < d[z]) d[z] = v; }
}
// generic 1-D min-plus pass along a strided line: out[i] = min_j in[j] + (i-j)^2
__global__ void edt_pass_line(const int* __restrict__ in, int* __restrict__ out, int S, size_t line_stride_a, size_t line_stride_b, size_t elem_stride)
{
    extern __shared__ int line[];
    const size_t base = (size_t)blockIdx.x * line_stride_a + (size_t)blockIdx.y * line_stride_b;
    for (int i = threadIdx.x; i < S; i += blockDim.x) line[i] = in[base + (size_t)i * elem_stride];
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        int best = kInf;
        for (int j = 0; j < S; j++) { const int v = line[j]; if (v < kInf) { const int c = v + (i - j) * (i - j); if (c < best) best = c; } }
        out[base + (size_t)i * elem_stride] = best;
    }
}
__global__ void edt_finalize_kernel(const int* __restrict__ D, int S, double scale, float* __restrict__ out)
{
    // D is [x][y][z]; out is [z][y][x]
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n3 = (size_t)S * S * S;
    if (i >= n3) return;
    const int x = (int)(i % S), y = (int)((i / S) % S), z = (int)(i / ((size_t)S * S));
    const int n2 = D[((size_t)x * S + y) * S + z];
    float dv = __double2float_rn(sqrt((double)n2));
    out[i] = __double2float_rn(__ddiv_rn((double)dv, scale));
}

} // namespace

void dt_frame_host(const float* m, int nm, int S, double expand, double* meta4)
{
    double xMin = m[0], xMax = m[0], yMin = m[1], yMax = m[1], zMin = m[2], zMax = m[2];
    for (int i = 1; i < nm; i++) {
        const double x = m[3 * i], y = m[3 * i + 1], z = m[3 * i + 2];
        if (xMin > x) xMin = x;
        if (xMax < x) xMax = x;
        if (yMin > y) yMin = y;
        if (yMax < y) yMax = y;
        if (zMin > z) zMin = z;
        if (zMax < z) zMax = z;
    }
    const double xc = (xMin + xMax) / 2, yc = (yMin + yMax) / 2, zc = (zMin + zMax) / 2;
    xMin = xc - expand * (xMax - xc); xMax = xc + expand * (xMax - xc);
    yMin = yc - expand * (yMax - yc); yMax = yc + expand * (yMax - yc);
    zMin = zc - expand * (zMax - zc); zMax = zc + expand * (zMax - zc);
    double side = xMax - xMin > yMax - yMin ? xMax - xMin : yMax - yMin;
    side = side > zMax - zMin ? side : zMax - zMin;
    meta4[0] = xc - side / 2; meta4[1] = yc - side / 2; meta4[2] = zc - side / 2;
    meta4[3] = S / side;
}

cudaError_t dt_build_device(const float* model, int nm, int S, double expand, int mode,
                            float* d_out, double* meta, cudaStream_t stream, std::string& msg)
{
    cudaError_t e;
    if (S > kMaxS) { msg = "dt_size > 1024"; return cudaErrorInvalidValue; }
    dt_frame_host(model, nm, S, expand, meta);
    const size_t n3 = (size_t)S * S * S;
    Vox* G = nullptr; float* d_model = nullptr; int* D0 = nullptr; int* D1 = nullptr;
    auto cleanup = [&]() { if (G) cudaFree(G); if (d_model) cudaFree(d_model); if (D0) cudaFree(D0); if (D1) cudaFree(D1); };
#define DT_TRY(expr) do { e = (expr); if (e != cudaSuccess) { msg = #expr; cleanup(); return e; } } while (0)
    DT_TRY(cudaMalloc((void**)&G, n3 * sizeof(Vox)));
    DT_TRY(cudaMalloc((void**)&d_model, (size_t)3 * nm * sizeof(float)));
    DT_TRY(cudaMemcpyAsync(d_model, model, (size_t)3 * nm * sizeof(float), cudaMemcpyHostToDevice, stream));
    dt_init_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, stream>>>(G, n3);
    DT_TRY(cudaGetLastError());
    dt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(G, S, d_model, nm, meta[0], meta[1], meta[2], meta[3]);
    DT_TRY(cudaGetLastError());
    if (mode == 0) {
        const int threads = ((S + 31) / 32) * 32;
        dt_propagate_kernel<<<1, threads, 0, stream>>>(G, S);
        DT_TRY(cudaGetLastError());
        dim3 grid((S + 31) / 32, (S + 31) / 32, S), block(32, 32);
        dt_finalize_kernel<<<grid, block, 0, stream>>>(G, S, meta[3], d_out);
        DT_TRY(cudaGetLastError());
    } else {
        // the extra corner seed is an artefact of the reference binary, not part of an exact EDT
        Vox unset; unset.v = unset.h = unset.d = 32767; unset.pad = 0;
        DT_TRY(cudaMemcpyAsync(G, &unset, sizeof(Vox), cudaMemcpyHostToDevice, stream));
        dt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(G, S, d_model, nm, meta[0], meta[1], meta[2], meta[3]);
        DT_TRY(cudaMalloc((void**)&D0, n3 * sizeof(int)));
        DT_TRY(cudaMalloc((void**)&D1, n3 * sizeof(int)));
        edt_pass_z<<<(S * S + 127) / 128, 128, 0, stream>>>(G, S, D0);
        DT_TRY(cudaGetLastError());
        // along y: lines indexed by (x, z): base = x*S*S + z, element stride S
        edt_pass_line<<<dim3(S, S), 128, S * sizeof(int), stream>>>(D0, D1, S, (size_t)S * S, 1, (size_t)S);
        DT_TRY(cudaGetLastError());
        // along x: lines indexed by (y, z): base = y*S + z, element stride S*S
        edt_pass_line<<<dim3(S, S), 128, S * sizeof(int), stream>>>(D1, D0, S, (size_t)S, 1, (size_t)S * S);
        DT_TRY(cudaGetLastError());
        edt_finalize_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, stream>>>(D0, S, meta[3], d_out);
        DT_TRY(cudaGetLastError());
    }
    DT_TRY(cudaStreamSynchronize(stream));
#undef DT_TRY
    cleanup();
    return cudaSuccess;
}

} // namespace goicp
