// Microbenchmark: issue/throughput of packed FFMA2 vs scalar FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    unsigned long long p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), ps = pk(s, s), pt = pk(0.5f, 0.25f);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {   // 8 independent scalar FFMA
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0 = fmaf(a0, s, 0.5f); a1 = fmaf(a1, s, 0.5f); a2 = fmaf(a2, s, 0.5f); a3 = fmaf(a3, s, 0.5f);
                a4 = fmaf(a4, s, 0.5f); a5 = fmaf(a5, s, 0.5f); a6 = fmaf(a6, s, 0.5f); a7 = fmaf(a7, s, 0.5f);
            }
        } else {           // 4 independent FFMA2 (= the same 8 FMAs)
#pragma unroll
            for (int j = 0; j < 8; ++j) { p0 = fma2(p0, ps, pt); p1 = fma2(p1, ps, pt); p2 = fma2(p2, ps, pt); p3 = fma2(p3, ps, pt); }
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(p0 ^ p1 ^ p2 ^ p3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000; const int blocks = 148 * 4, threads = 512;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(d, iters, 0.999f); else k<1><<<blocks, threads>>>(d, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)blocks * threads * iters * 64.0;
            if (rep == 2) printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode == 0 ? "FFMA " : "FFMA2", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
        }
    }
    return 0;
}
