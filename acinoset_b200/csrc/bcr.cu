// Block cyclic reduction of the block-tridiagonal LM normal equations (fp64, 75x75 super-blocks).
//
// Replaces the linear algebra inside the reference's `opt.solve(m)` IPOPT call
// (/root/reference/src/all_optimizations.py:503-524; IPOPT's default MUMPS factorisation of the
// KKT system) for the reduced problem of SURVEY.md appendix B6.  Host-side schedule:
// acinoset_b200/bcr.py.  Per level three kernels, one CTA per super-block:
//   bcr_factor   (eliminated block e, neighbours a/c):  D_e = R R^T,
//                P = R^-1 Lc_e,  Q = R^-1 Lc_c^T,  z = R^-1 b_e   (one fused elimination; R kept for back-subst)
//   bcr_update   (surviving block j, eliminated neighbours el/er):
//                D_j -= Q_el^T Q_el + P_er^T P_er,  b_j -= Q_el^T z_el + P_er^T z_er,
//                Lc_j = -Q_el^T P_el
//   bcr_backsub  (reverse order)   x_e = R^-T (z - P x_a - Q x_c)
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

// ---- bcr_factor: fused elimination of one super-block -------------------------------------------
// Gaussian elimination (no pivoting; the block is SPD) of the augmented panel
//     [ D_e | Lc_e | Lc_c^T | b_e ]      75 x (75 + 75 + 75 + 1)
// held COLUMN-major in shared memory (column stride FCS = 76 doubles: 16-byte accesses of a quarter-warp
// = 4 columns x 2 row pairs are bank-conflict-free, the pivot column is a broadcast).
// Two threads own column j (interleaved row pairs).  Step k: a_ij -= a_ik (a_kj / a_kk) for i > k (D columns: i >= j only, the
// row-k entry a_kj is read from the lower triangle as a_jk).  After step k-1 row k of the right-hand
// columns is (L^-1 [Lc_e | Lc_c^T | b])_k with D = L Delta L^T, so
//     P = R^-1 Lc_e, Q = R^-1 Lc_c^T, z = R^-1 b   with R = L Delta^1/2   are row scalings by a_kk^-1/2.
// One barrier per step; the owner of the next pivot column updates its diagonal entry first and
// publishes 1/a_kk.  No explicit inverse, no separate GEMMs: 0.5 MFMA per block instead of 1.1.
// Stored for back-substitution in place of D_e: strict lower triangle = L (unit diagonal implied),
// diagonal = a_kk^-1/2.
constexpr int FCOLS = 3 * SB + 1;   // 226 panel columns
constexpr int FCS = 76;             // column stride (doubles): 152 words = 24 (mod 32) -> 4 columns x 2 row-pairs tile the banks
constexpr int FACTOR_THREADS = 512; // two threads per column (interleaved row pairs)
constexpr size_t BCR_FACTOR_SMEM = (size_t)FCOLS * FCS * sizeof(double);

__global__ void __launch_bounds__(FACTOR_THREADS)
bcr_factor_kernel(const int* __restrict__ elim /*[ne][3]*/, double* __restrict__ D, const double* __restrict__ Lc,
                  double* __restrict__ P, double* __restrict__ Q, double* __restrict__ rhs, int* __restrict__ info) {
    extern __shared__ __align__(16) double sm[];     // [FCOLS][FCS]
    __shared__ double sinv[SB], sdi[SB];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    // ---- load the panel (column-major); row 75 of every column is zero padding
    for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
        const int r = i / SB, cc = i - r * SB;
        sm[r * FCS + cc] = D[(size_t)e * SB2 + i];                                      // symmetric: (r,cc) == (cc,r)
        if (a >= 0) sm[(SB + cc) * FCS + r] = Lc[(size_t)e * SB2 + i];                  // Lc_e (r, cc)
        if (c >= 0) sm[(2 * SB + r) * FCS + cc] = Lc[(size_t)c * SB2 + i];              // Lc_c^T (cc, r) = Lc_c (r, cc)
    }
    if (tid < SB) sm[3 * SB * FCS + tid] = rhs[(size_t)e * SB + tid];
    if (tid < FCOLS) sm[tid * FCS + SB] = 0.0;
    __syncthreads();
    if (tid == 0) {
        const double d = sm[0];
        if (!(d > 0.0)) atomicExch(info, e + 1);
        sinv[0] = 1.0 / d;
    }
    const int col = tid >> 1, half = tid & 1;
    const bool is_d = col < SB;
    const bool active_col = col < FCOLS && (is_d || col >= 3 * SB || (col < 2 * SB ? a >= 0 : c >= 0));
    double* cj = sm + col * FCS;
    for (int k = 0; k < SB - 1; ++k) {
        __syncthreads();
        if (!active_col || col <= k) continue;
        const double* ck = sm + k * FCS;
        const double m = -((is_d ? ck[col] : cj[k]) * sinv[k]);
        int i = is_d ? col : k + 1;              // first row this column updates
        if (col == k + 1 && half == 0) {         // owner of the next pivot: finish a_(k+1)(k+1), publish its inverse
            const double d = fma(ck[i], m, cj[i]);
            cj[i] = d;
            if (!(d > 0.0)) atomicExch(info, e + 1);
            sinv[k + 1] = 1.0 / d;
        }
        if (col == k + 1) ++i;
        if (i & 1) {                             // odd first row: a single-row update, then aligned pairs
            if (half == 0) cj[i] = fma(ck[i], m, cj[i]);
            ++i;
        }
        // aligned row pairs (i, i+1), interleaved over the column's two threads; the last pair touches padding
        i += 2 * half;
        for (; i + 12 < SB + 1; i += 16) {       // four pairs per trip: all loads, then the FMAs, then the stores
            const double2 p0 = *reinterpret_cast<const double2*>(ck + i), p1 = *reinterpret_cast<const double2*>(ck + i + 4);
            const double2 p2 = *reinterpret_cast<const double2*>(ck + i + 8), p3 = *reinterpret_cast<const double2*>(ck + i + 12);
            double2 q0 = *reinterpret_cast<double2*>(cj + i), q1 = *reinterpret_cast<double2*>(cj + i + 4);
            double2 q2 = *reinterpret_cast<double2*>(cj + i + 8), q3 = *reinterpret_cast<double2*>(cj + i + 12);
            q0.x = fma(p0.x, m, q0.x); q0.y = fma(p0.y, m, q0.y);
            q1.x = fma(p1.x, m, q1.x); q1.y = fma(p1.y, m, q1.y);
            q2.x = fma(p2.x, m, q2.x); q2.y = fma(p2.y, m, q2.y);
            q3.x = fma(p3.x, m, q3.x); q3.y = fma(p3.y, m, q3.y);
            *reinterpret_cast<double2*>(cj + i) = q0;
            *reinterpret_cast<double2*>(cj + i + 4) = q1;
            *reinterpret_cast<double2*>(cj + i + 8) = q2;
            *reinterpret_cast<double2*>(cj + i + 12) = q3;
        }
        for (; i < SB; i += 4) {
            const double2 p0 = *reinterpret_cast<const double2*>(ck + i);
            double2 q0 = *reinterpret_cast<double2*>(cj + i);
            q0.x = fma(p0.x, m, q0.x);
            q0.y = fma(p0.y, m, q0.y);
            *reinterpret_cast<double2*>(cj + i) = q0;
        }
    }
    __syncthreads();
    if (tid < SB) sdi[tid] = sqrt(fmax(sinv[tid], 0.0));
    __syncthreads();
    // ---- write-out: factor in place of D_e, P, Q, z
    for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
        const int r = i / SB, cc = i - r * SB;
        D[(size_t)e * SB2 + i] = r > cc ? sm[cc * FCS + r] * sinv[cc] : (r == cc ? sdi[r] : 0.0);
        if (a >= 0) P[(size_t)e * SB2 + i] = sm[(SB + cc) * FCS + r] * sdi[r];
        if (c >= 0) Q[(size_t)e * SB2 + i] = sm[(2 * SB + cc) * FCS + r] * sdi[r];
    }
    if (tid < SB) rhs[(size_t)e * SB + tid] = sm[3 * SB * FCS + tid] * sdi[tid];
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

// x_e = R^-T (z - P x_a - Q x_c), R = L Delta^1/2 stored by bcr_factor in D_e.  256 threads.
//   1. v = z - [P | Q] [x_a ; x_c]: thread (row, third) reads 50 contiguous doubles of the row's 150 - all loads
//      independent and issued at once; L is staged in shared memory meanwhile
//   2. backward substitution L^T x = Delta^-1/2 v by ONE warp, no barriers: lane holds rows lane, lane+32,
//      lane+64; once x_k is final it is broadcast by shuffle and u_i -= L_ki x_k for i < k (row k of L).
__global__ void __launch_bounds__(256)
bcr_backsub_kernel(const int* __restrict__ elim, const double* __restrict__ D /*factor*/, const double* __restrict__ P,
                   const double* __restrict__ Q, const double* __restrict__ rhs, double* __restrict__ x) {
    __shared__ __align__(16) double sxx[2 * SB + 2];       // [x_a ; x_c]
    __shared__ double spart[3][SB];
    __shared__ double sL[SB * LD];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    if (tid < SB) sxx[tid] = a >= 0 ? x[(size_t)a * SB + tid] : 0.0;
    else if (tid < 2 * SB) sxx[tid] = c >= 0 ? x[(size_t)c * SB + tid - SB] : 0.0;
    for (int i = tid; i < SB * SB; i += 256) {
        const int r = i / SB;
        sL[r * LD + (i - r * SB)] = D[(size_t)e * SB2 + i];
    }
    __syncthreads();
    if (tid < 3 * SB) {
        const int r = tid / 3, part = tid - 3 * r;          // columns [50 part, 50 part + 50) of [P | Q]
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 50; k += 2) {
            const int cc = 50 * part + k;                    // even; a pair never straddles P | Q (75 is odd: handle per element)
            const int c0 = cc, c1 = cc + 1;
            const double m0 = c0 < SB ? (a >= 0 ? P[(size_t)e * SB2 + r * SB + c0] : 0.0)
                                      : (c >= 0 ? Q[(size_t)e * SB2 + r * SB + c0 - SB] : 0.0);
            const double m1 = c1 < SB ? (a >= 0 ? P[(size_t)e * SB2 + r * SB + c1] : 0.0)
                                      : (c >= 0 ? Q[(size_t)e * SB2 + r * SB + c1 - SB] : 0.0);
            s0 = fma(m0, sxx[c0], s0);
            s1 = fma(m1, sxx[c1], s1);
        }
        spart[part][r] = s0 + s1;
    }
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        double u[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = lane + 32 * j;
            u[j] = i < SB ? (rhs[(size_t)e * SB + i] - ((spart[0][i] + spart[1][i]) + spart[2][i])) * sL[i * LD + i] : 0.0;
        }
#pragma unroll
        for (int j = 2; j >= 0; --j) {
            for (int kk = (j == 2 ? SB - 1 - 64 : 31); kk >= 0; --kk) {
                const int k = 32 * j + kk;
                const double xk = __shfl_sync(0xffffffffu, u[j], kk);
                // rows i < k: slots below j entirely, slot j for lanes < kk
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    if (jj > j) continue;
                    const int i = lane + 32 * jj;
                    if (jj < j || lane < kk) u[jj] = fma(-sL[k * LD + i], xk, u[jj]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = lane + 32 * j;
            if (i < SB) x[(size_t)e * SB + i] = u[j];
        }
    }
}

constexpr size_t BCR_SMEM = 2 * SB * LD * sizeof(double);

cudaError_t launch_bcr_factor(int n_elim, const int* elim, double* D, const double* Lc, double* P, double* Q,
                              double* rhs, int* info, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(bcr_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_FACTOR_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(bcr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_SMEM);
        if (e != cudaSuccess) return e;
        set = true;
    }
    bcr_factor_kernel<<<n_elim, FACTOR_THREADS, BCR_FACTOR_SMEM, s>>>(elim, D, Lc, P, Q, rhs, info);
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
    bcr_backsub_kernel<<<n_elim, 256, 0, s>>>(elim, D, P, Q, rhs, x);
    return cudaGetLastError();
}

}  // namespace acino
