// Generic-skeleton FTE kernels (the data-driven variant of the reference, /root/reference/src/build.py:28-332), fp64.
// The arithmetic lives in skel_body.cuh (shared with the CPU test harness); this file binds it to CUDA launches.
//   skel_eval      one CTA per frame: {cost, g[P], H[P(P+1)/2]} of the measurement term
//   skel_prepare   total gradient, frozen (bound-active) flags, per-frame smoothness cost
//   skel_assemble  lower band of (B + lam diag B), B = blockdiag(H_n) + sw D3^T D3 (half bandwidth 3P)
//   band_solve     in-place band Cholesky + substitution, one CTA
//   skel_trial / skel_pred   projected trial point, quadratic-model reduction and step norm per frame
// The human-skeleton problems the reference ships are 100 frames x 48 parameters (build.py:133-135); these kernels
// are the correct, table-driven path for ANY skeleton pickle, not tuned like the cheetah kernels (fte_eval.cu).
#include <cuda_runtime.h>

#include "skel_body.cuh"

namespace acino {

struct CtaCtx {
    int tid, nthreads;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // barrier among the first n threads of the CTA (n a multiple of 32), named barrier 1
    __device__ __forceinline__ void sync_part(int n) const { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
};
struct GridCtx {   // grid-stride phases without barriers
    long long tid, nthreads;
    __device__ __forceinline__ void sync() const {}
    __device__ __forceinline__ void sync_part(int) const {}
};

__global__ void __launch_bounds__(128)
skel_eval_kernel(const SkelDesc* __restrict__ S, const int P, const int n_frames, const double* __restrict__ x,
                 const double* __restrict__ meas, const double* __restrict__ w, double* __restrict__ cost,
                 double* __restrict__ g, double* __restrict__ H) {
    extern __shared__ double sk_sm[];
    const int n = blockIdx.x;
    if (n >= n_frames) return;
    const CtaCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    const size_t mo = (size_t)S->n_cams * S->n_out;
    skel_eval_frame(*S, ctx, x + (size_t)n * P, meas + (size_t)n * mo * 2, w + (size_t)n * mo, cost ? cost + n : nullptr,
                    g ? g + (size_t)n * P : nullptr, H ? H + (size_t)n * (P * (P + 1) / 2) : nullptr, sk_sm);
}

__global__ void skel_prepare_kernel(int N, int P, int last_free, const double* x, const double* g, const double* sw,
                                    const double* lo, const double* hi, double* gtot, unsigned char* fixed, double* cost_s) {
    const GridCtx ctx{(long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x};
    skel_prepare(ctx, N, P, last_free, x, g, sw, lo, hi, gtot, fixed, cost_s);
}

__global__ void skel_assemble_kernel(int N, int P, const double* H, const double* gtot, const unsigned char* fixed,
                                     const double* sw, double lam, double* AB, double* rhs) {
    const GridCtx ctx{(long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x};
    skel_assemble(ctx, N, P, H, gtot, fixed, sw, lam, AB, rhs);
}

constexpr int BAND_NB = 16;     // panel width of the blocked band Cholesky
__global__ void __launch_bounds__(1024) band_solve_kernel(long long n, int hb, double* AB, double* x, int* info) {
    extern __shared__ double band_sm[];
    const CtaCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    band_cholesky_solve<BAND_NB>(ctx, n, hb, AB, x, info, band_sm);
}

__global__ void skel_trial_kernel(int N, int P, int last_free, const double* x, const double* d, const double* lo,
                                  const double* hi, double* xt) {
    const GridCtx ctx{(long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x};
    skel_trial(ctx, N, P, last_free, x, d, lo, hi, xt);
}

__global__ void skel_pred_kernel(int N, int P, const double* x, const double* xt, const double* gtot, const double* H,
                                 const double* sw, double* pred, double* step) {
    const GridCtx ctx{(long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x};
    skel_pred(ctx, N, P, x, xt, gtot, H, sw, pred, step);
}

static inline int grid_for(long long n, int b) {
    long long g = (n + b - 1) / b;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

size_t skel_eval_smem_bytes(int n_links, int n_out) { return (size_t)SkelSmemLayout(n_links, n_out).total * sizeof(double); }

cudaError_t launch_skel_eval(const SkelDesc* d_skel, int n_links, int n_out, int P, int n_frames, const double* x,
                             const double* meas, const double* w, double* cost, double* g, double* H, cudaStream_t s) {
    if (n_frames <= 0) return cudaSuccess;
    const size_t smem = skel_eval_smem_bytes(n_links, n_out);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(skel_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    skel_eval_kernel<<<n_frames, 128, smem, s>>>(d_skel, P, n_frames, x, meas, w, cost, g, H);
    return cudaGetLastError();
}
cudaError_t launch_skel_prepare(int N, int P, int last_free, const double* x, const double* g, const double* sw,
                                const double* lo, const double* hi, double* gtot, unsigned char* fixed, double* cost_s,
                                cudaStream_t s) {
    skel_prepare_kernel<<<grid_for((long long)N * P, 256), 256, 0, s>>>(N, P, last_free, x, g, sw, lo, hi, gtot, fixed, cost_s);
    return cudaGetLastError();
}
cudaError_t launch_skel_assemble(int N, int P, const double* H, const double* gtot, const unsigned char* fixed,
                                 const double* sw, double lam, double* AB, double* rhs, cudaStream_t s) {
    skel_assemble_kernel<<<grid_for((long long)N * P * (3 * P + 1), 256), 256, 0, s>>>(N, P, H, gtot, fixed, sw, lam, AB, rhs);
    return cudaGetLastError();
}
cudaError_t launch_band_solve(long long n, int hb, double* AB, double* x, int* info, cudaStream_t s) {
    const size_t smem = band_panel_doubles(hb, BAND_NB) * sizeof(double);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(band_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    band_solve_kernel<<<1, 1024, smem, s>>>(n, hb, AB, x, info);
    return cudaGetLastError();
}
cudaError_t launch_skel_trial(int N, int P, int last_free, const double* x, const double* d, const double* lo,
                              const double* hi, double* xt, cudaStream_t s) {
    skel_trial_kernel<<<grid_for((long long)N * P, 256), 256, 0, s>>>(N, P, last_free, x, d, lo, hi, xt);
    return cudaGetLastError();
}
cudaError_t launch_skel_pred(int N, int P, const double* x, const double* xt, const double* gtot, const double* H,
                             const double* sw, double* pred, double* step, cudaStream_t s) {
    skel_pred_kernel<<<grid_for(N, 64), 64, 0, s>>>(N, P, x, xt, gtot, H, sw, pred, step);
    return cudaGetLastError();
}

}  // namespace acino
