// Dependent-issue latency of the float add the reference-order sums are made of, one warp (and with 15 idle-looping
// neighbours): (a) operands in registers, FADD; (b) same, FFMA(x, 1.0f, acc) -- same result, one rounding; (c) operands streamed
// from shared memory 16 ahead, the loop of seq_add_contig (icp_kernels.cu).      nvcc -arch=sm_100a -fmad=false -o dep_chain dep_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void chain(const float* __restrict__ x, float* out, long long* cyc, int n)
{
    __shared__ __align__(16) float sx[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    float acc = 0.0f;
    long long t0 = 0, t1 = 0;
    if (MODE < 2) {
        float c[16];
#pragma unroll
        for (int k = 0; k < 16; k++) c[k] = sx[k + threadIdx.x];
        t0 = clock64();
        for (int rep = 0; rep < n * 256; rep++) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc = MODE == 0 ? __fadd_rn(acc, c[k]) : __fmaf_rn(c[k], 1.0f, acc);
        }
        t1 = clock64();
    } else {
        const float4* x4 = reinterpret_cast<const float4*>(sx);
        t0 = clock64();
        for (int rep = 0; rep < n; rep++) {
            float4 a0 = x4[0], a1 = x4[1], a2 = x4[2], a3 = x4[3];
            for (int rr = 16; rr + 16 <= 4096; rr += 16) {
                const float4 b0 = x4[rr / 4], b1 = x4[rr / 4 + 1], b2 = x4[rr / 4 + 2], b3 = x4[rr / 4 + 3];
                acc = __fadd_rn(acc, a0.x); acc = __fadd_rn(acc, a0.y); acc = __fadd_rn(acc, a0.z); acc = __fadd_rn(acc, a0.w);
                acc = __fadd_rn(acc, a1.x); acc = __fadd_rn(acc, a1.y); acc = __fadd_rn(acc, a1.z); acc = __fadd_rn(acc, a1.w);
                acc = __fadd_rn(acc, a2.x); acc = __fadd_rn(acc, a2.y); acc = __fadd_rn(acc, a2.z); acc = __fadd_rn(acc, a2.w);
                acc = __fadd_rn(acc, a3.x); acc = __fadd_rn(acc, a3.y); acc = __fadd_rn(acc, a3.z); acc = __fadd_rn(acc, a3.w);
                a0 = b0; a1 = b1; a2 = b2; a3 = b3;
            }
            acc = __fadd_rn(acc, a0.x); acc = __fadd_rn(acc, a0.y); acc = __fadd_rn(acc, a0.z); acc = __fadd_rn(acc, a0.w);
            acc = __fadd_rn(acc, a1.x); acc = __fadd_rn(acc, a1.y); acc = __fadd_rn(acc, a1.z); acc = __fadd_rn(acc, a1.w);
            acc = __fadd_rn(acc, a2.x); acc = __fadd_rn(acc, a2.y); acc = __fadd_rn(acc, a2.z); acc = __fadd_rn(acc, a2.w);
            acc = __fadd_rn(acc, a3.x); acc = __fadd_rn(acc, a3.y); acc = __fadd_rn(acc, a3.z); acc = __fadd_rn(acc, a3.w);
        }
        t1 = clock64();
    }
    if (threadIdx.x == 0) { out[MODE] = acc; cyc[MODE] = t1 - t0; }
}
int main()
{
    float* x; float* out; long long* cyc;
    cudaMallocManaged(&x, 4096 * 4 + 256); cudaMallocManaged(&out, 16); cudaMallocManaged(&cyc, 32);
    for (int i = 0; i < 4096 + 64; i++) x[i] = 1.0f / (float)(i + 3);
    for (int pass = 0; pass < 2; pass++) {
        chain<0><<<1, 512>>>(x, out, cyc, 8); chain<1><<<1, 512>>>(x, out, cyc, 8); chain<2><<<1, 512>>>(x, out, cyc, 8);
        cudaDeviceSynchronize();
    }
    printf("registers: FADD %.2f cycles/add, FFMA(x,1,acc) %.2f cycles/add (sums %s); shared-memory stream (seq_add_contig): %.2f cycles/add\n",
           cyc[0] / (8.0 * 4096), cyc[1] / (8.0 * 4096), out[0] == out[1] ? "equal" : "DIFFER", cyc[2] / (8.0 * 4096));
    return 0;
}
