// Dependent-chain latency of FFMA vs FFMA2 (single warp per SM, clock64)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, long long* cyc, int iters, float s) {
    float a = threadIdx.x; unsigned long long p = pk(a, a + 1), ps = pk(s, s * 1.01f), pt = pk(0.5f, 0.25f);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { if (MODE == 0) a = fmaf(a, s, 0.5f); else p = fma2(p, ps, pt); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + (float)p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    for (int mode = 0; mode < 2; ++mode) {
        const int iters = 1000; long long h;
        for (int rep = 0; rep < 2; ++rep) { if (mode == 0) k<0><<<1, 32>>>(d, c, iters, 0.999f); else k<1><<<1, 32>>>(d, c, iters, 0.999f); }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%s dependent latency: %.2f cycles\n", mode == 0 ? "FFMA " : "FFMA2", (double)h / (iters * 32.0));
    }
    return 0;
}
