// Block cyclic reduction of the block-tridiagonal LM normal equations (fp64, 75x75 super-blocks).
//
// Replaces the linear algebra inside the reference's `opt.solve(m)` IPOPT call
// (/root/reference/src/all_optimizations.py:503-524; IPOPT's default MUMPS factorisation of the
// KKT system) for the reduced problem of SURVEY.md appendix B6.  Host-side schedule:
// acinoset_b200/bcr.py.  Per level three kernels, one CTA per super-block:
//   bcr_factor   (eliminated block e, neighbours a/c):  D_e = R R^T,  W = R^-1,
//                P = W Lc_e,  Q = W Lc_c^T,  z = W b_e                      (stored for back-subst)
//   bcr_update   (surviving block j, eliminated neighbours el/er):
//                D_j -= Q_el^T Q_el + P_er^T P_er,  b_j -= Q_el^T z_el + P_er^T z_er,
//                Lc_j = -Q_el^T P_el
//   bcr_backsub  (reverse order)   x_e = W^T (z - P x_a - Q x_c)
// fp64 throughout: the smoothness weights (2 q / Ts^4 ~ 1e7..1e9) against data blocks (~1e5) make
// the system too ill-conditioned for fp32 factorisation; B200 runs fp64 FMA at half the fp32 rate.
#include "acino_common.cuh"

namespace acino {

constexpr int SB = 75;          // super-block size: 3 frames x 25 parameters
constexpr int LD = 76;          // shared-memory leading dimension (doubles)
constexpr int BCR_THREADS = 256;
constexpr size_t SB2 = (size_t)SB * SB;

// C(75x75 tile set) = op(A) * B with A, B in shared memory (ld = LD); each thread owns a 5x5 tile:
// ty = tid / 16 -> rows 5 ty .. 5 ty + 4, tx = tid % 16 -> cols 5 tx .. 5 tx + 4 (80 x 80 cover).
template <bool TRANS_A>
__device__ __forceinline__ void gemm75(const double* __restrict__ sA, const double* __restrict__ sB, double acc[5][5]) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int i0 = 5 * ty, j0 = 5 * tx;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[r][c] = 0.0;
    if (i0 >= SB || j0 >= SB) return;
    for (int k = 0; k < SB; ++k) {
        double a[5], b[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) a[r] = TRANS_A ? sA[k * LD + i0 + r] : sA[(i0 + r) * LD + k];
#pragma unroll
        for (int c = 0; c < 5; ++c) b[c] = sB[k * LD + j0 + c];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
}

// global (row-major 75x75) -> shared (ld 76); optionally transposed
__device__ __forceinline__ void load_block(const double* __restrict__ g, double* __restrict__ s, bool transpose) {
    for (int i = threadIdx.x; i < SB * SB; i += BCR_THREADS) {
        const int r = i / SB, c = i - r * SB;
        if (transpose) s[c * LD + r] = g[i];
        else s[r * LD + c] = g[i];
    }
}

__global__ void __launch_bounds__(BCR_THREADS)
bcr_factor_kernel(const int* __restrict__ elim /*[ne][3]*/, double* __restrict__ D, const double* __restrict__ Lc,
                  double* __restrict__ P, double* __restrict__ Q, double* __restrict__ rhs, int* __restrict__ info) {
    extern __shared__ __align__(16) double sm[];
    double* sR = sm;                 // Cholesky workspace, later the right-hand-side panel
    double* sW = sm + SB * LD;       // W = R^-1 (lower triangular, zeros above)
    __shared__ double sz[SB], sz2[SB];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    load_block(D + (size_t)e * SB2, sR, false);
    for (int i = tid; i < SB * LD; i += BCR_THREADS) sW[i] = 0.0;
    __syncthreads();
    // right-looking Cholesky, one barrier per column: a_ij -= a_ik a_jk / a_kk  (i >= j > k).
    // 16 x 16 thread grid strided over the trailing matrix (no index division in the loop)
    {
        const int ty = tid >> 4, tx = tid & 15;
        for (int k = 0; k < SB - 1; ++k) {
            const double inv = 1.0 / sR[k * LD + k];
            for (int i = k + 1 + ty; i < SB; i += 16) {
                const double lik = -sR[i * LD + k] * inv;
                for (int j = k + 1 + tx; j <= i; j += 16) sR[i * LD + j] = fma(lik, sR[j * LD + k], sR[i * LD + j]);
            }
            __syncthreads();
        }
    }
    // scale columns: R_ik = a_ik / sqrt(a_kk); flag non-positive pivots
    for (int t = tid; t < SB * SB; t += BCR_THREADS) {
        const int i = t / SB, k = t - i * SB;
        if (i >= k) {
            const double d = sR[k * LD + k];
            if (i == k && !(d > 0.0)) atomicExch(info, e + 1);
        }
    }
    __syncthreads();
    if (tid < SB) sz[tid] = sqrt(fmax(sR[tid * LD + tid], 1e-300));
    __syncthreads();
    for (int t = tid; t < SB * SB; t += BCR_THREADS) {
        const int i = t / SB, k = t - i * SB;
        if (i > k) sR[i * LD + k] = sR[i * LD + k] / sz[k];
    }
    __syncthreads();
    if (tid < SB) {
        sR[tid * LD + tid] = sz[tid];
        sz2[tid] = 1.0 / sz[tid];
    }
    __syncthreads();
    // W = R^-1: thread j solves R w = e_j by forward substitution (column j of W)
    if (tid < SB) {
        const int j = tid;
        sW[j * LD + j] = sz2[j];
        for (int i = j + 1; i < SB; ++i) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = j;
            for (; k + 3 < i; k += 4) {
                s0 = fma(sR[i * LD + k], sW[k * LD + j], s0);
                s1 = fma(sR[i * LD + k + 1], sW[(k + 1) * LD + j], s1);
                s2 = fma(sR[i * LD + k + 2], sW[(k + 2) * LD + j], s2);
                s3 = fma(sR[i * LD + k + 3], sW[(k + 3) * LD + j], s3);
            }
            for (; k < i; ++k) s0 = fma(sR[i * LD + k], sW[k * LD + j], s0);
            sW[i * LD + j] = -((s0 + s1) + (s2 + s3)) * sz2[i];
        }
    }
    __syncthreads();
    // store W in place of D_e
    for (int i = tid; i < SB * SB; i += BCR_THREADS) {
        const int r = i / SB, cc = i - r * SB;
        D[(size_t)e * SB2 + i] = sW[r * LD + cc];
    }
    // z = W b_e
    if (tid < SB) sz2[tid] = rhs[(size_t)e * SB + tid];
    __syncthreads();
    if (tid < SB) {
        double s = 0.0;
        for (int k = 0; k <= tid; ++k) s = fma(sW[tid * LD + k], sz2[k], s);
        rhs[(size_t)e * SB + tid] = s;
    }
    const int ty = tid >> 4, tx = tid & 15;
    double acc[5][5];
    // P = W Lc_e
    if (a >= 0) {
        __syncthreads();
        load_block(Lc + (size_t)e * SB2, sR, false);
        __syncthreads();
        gemm75<false>(sW, sR, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int cc = 0; cc < 5; ++cc) {
                const int i = 5 * ty + r, j = 5 * tx + cc;
                if (i < SB && j < SB) P[(size_t)e * SB2 + i * SB + j] = acc[r][cc];
            }
    }
    // Q = W Lc_c^T
    if (c >= 0) {
        __syncthreads();
        load_block(Lc + (size_t)c * SB2, sR, true);
        __syncthreads();
        gemm75<false>(sW, sR, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int cc = 0; cc < 5; ++cc) {
                const int i = 5 * ty + r, j = 5 * tx + cc;
                if (i < SB && j < SB) Q[(size_t)e * SB2 + i * SB + j] = acc[r][cc];
            }
    }
}

__global__ void __launch_bounds__(BCR_THREADS)
bcr_update_kernel(const int* __restrict__ surv /*[ns][3]*/, double* __restrict__ D, double* __restrict__ Lc,
                  const double* __restrict__ P, const double* __restrict__ Q, double* __restrict__ rhs) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;
    double* sB = sm + SB * LD;
    __shared__ double sz[SB];
    const int j = surv[3 * blockIdx.x], el = surv[3 * blockIdx.x + 1], er = surv[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    double accD[5][5], acc[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) accD[r][c] = 0.0;
    double rj = 0.0;
    if (el >= 0) {
        load_block(Q + (size_t)el * SB2, sA, false);
        load_block(P + (size_t)el * SB2, sB, false);
        if (tid < SB) sz[tid] = rhs[(size_t)el * SB + tid];
        __syncthreads();
        // Lc_j = -Q^T P
        gemm75<true>(sA, sB, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const int i = 5 * ty + r, jj = 5 * tx + c;
                if (i < SB && jj < SB) Lc[(size_t)j * SB2 + i * SB + jj] = -acc[r][c];
            }
        // D_j -= Q^T Q ; b_j -= Q^T z
        gemm75<true>(sA, sA, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) accD[r][c] += acc[r][c];
        if (tid < SB) {
            double s = 0.0;
            for (int k = 0; k < SB; ++k) s = fma(sA[k * LD + tid], sz[k], s);
            rj += s;
        }
        __syncthreads();
    }
    if (er >= 0) {
        load_block(P + (size_t)er * SB2, sA, false);
        if (tid < SB) sz[tid] = rhs[(size_t)er * SB + tid];
        __syncthreads();
        gemm75<true>(sA, sA, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) accD[r][c] += acc[r][c];
        if (tid < SB) {
            double s = 0.0;
            for (int k = 0; k < SB; ++k) s = fma(sA[k * LD + tid], sz[k], s);
            rj += s;
        }
    }
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int i = 5 * ty + r, jj = 5 * tx + c;
            if (i < SB && jj < SB) D[(size_t)j * SB2 + i * SB + jj] -= accD[r][c];
        }
    if (tid < SB) rhs[(size_t)j * SB + tid] -= rj;
}

// x_e = W^T (z - P x_a - Q x_c); 128 threads, one warp per row for the two mat-vecs
__global__ void __launch_bounds__(128)
bcr_backsub_kernel(const int* __restrict__ elim, const double* __restrict__ D /*W*/, const double* __restrict__ P,
                   const double* __restrict__ Q, const double* __restrict__ rhs, double* __restrict__ x) {
    __shared__ double sxa[SB], sxc[SB], sv[SB];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < SB) {
        sxa[tid] = a >= 0 ? x[(size_t)a * SB + tid] : 0.0;
        sxc[tid] = c >= 0 ? x[(size_t)c * SB + tid] : 0.0;
    }
    __syncthreads();
    for (int i = warp; i < SB; i += 4) {
        double s = 0.0;
        if (a >= 0)
            for (int k = lane; k < SB; k += 32) s = fma(P[(size_t)e * SB2 + i * SB + k], sxa[k], s);
        if (c >= 0)
            for (int k = lane; k < SB; k += 32) s = fma(Q[(size_t)e * SB2 + i * SB + k], sxc[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sv[i] = rhs[(size_t)e * SB + i] - s;
    }
    __syncthreads();
    if (tid < SB) {
        double s = 0.0;
        for (int i = tid; i < SB; ++i) s = fma(D[(size_t)e * SB2 + i * SB + tid], sv[i], s);   // W lower-triangular
        x[(size_t)e * SB + tid] = s;
    }
}

constexpr size_t BCR_SMEM = 2 * SB * LD * sizeof(double);

cudaError_t launch_bcr_factor(int n_elim, const int* elim, double* D, const double* Lc, double* P, double* Q,
                              double* rhs, int* info, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(bcr_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(bcr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_SMEM);
        if (e != cudaSuccess) return e;
        set = true;
    }
    bcr_factor_kernel<<<n_elim, BCR_THREADS, BCR_SMEM, s>>>(elim, D, Lc, P, Q, rhs, info);
    return cudaGetLastError();
}

cudaError_t launch_bcr_update(int n_surv, const int* surv, double* D, double* Lc, const double* P, const double* Q,
                              double* rhs, cudaStream_t s) {
    if (n_surv <= 0) return cudaSuccess;
    bcr_update_kernel<<<n_surv, BCR_THREADS, BCR_SMEM, s>>>(surv, D, Lc, P, Q, rhs);
    return cudaGetLastError();
}

cudaError_t launch_bcr_backsub(int n_elim, const int* elim, const double* D, const double* P, const double* Q,
                               const double* rhs, double* x, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    bcr_backsub_kernel<<<n_elim, 128, 0, s>>>(elim, D, P, Q, rhs, x);
    return cudaGetLastError();
}

}  // namespace acino
