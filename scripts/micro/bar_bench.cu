// microbenchmark: cost of one CTA-wide barrier step as a function of warp count, and of the
// shared-memory exchange pattern of the DT propagation (STS.128 -> bar -> 3x LDS.128)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void bar_loop(int iters, long long* out, int mode)
{
    extern __shared__ int4 buf[];
    int4 v = make_int4(threadIdx.x, 1, 2, 3);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (mode >= 1) buf[(i & 1) * 1100 + threadIdx.x + 1] = v;
        asm volatile("bar.sync 0;" ::: "memory");
        if (mode >= 1) {
            int4 a = buf[(i & 1) * 1100 + threadIdx.x], b = buf[(i & 1) * 1100 + threadIdx.x + 2];
            v.x += a.x + b.y; v.y ^= a.z; v.z += b.w; v.w += a.w;
        }
        if (mode >= 2) {   // a shuffle + vote round as in the scans
            int q = __shfl_up_sync(0xffffffffu, v.x, 1), r = __shfl_down_sync(0xffffffffu, v.y, 1);
            if (__any_sync(0xffffffffu, (q + r) == 123456789)) v.x++;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    if (v.x == 0x7fffffff) out[1] = v.y + v.z + v.w;
}
int main()
{
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bar_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 1100 * 16);
    for (int mode = 0; mode < 3; mode++)
        for (int warps : {1, 2, 4, 5, 10, 11, 21, 32}) {
            const int iters = 20000;
            bar_loop<<<1, warps * 32, 2 * 1100 * 16>>>(iters, d, mode);
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode %d warps %2d: %.1f cycles/step\n", mode, warps, (double)h[0] / iters);
        }
    return 0;
}
