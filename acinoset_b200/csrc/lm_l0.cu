// Structured first level of the block cyclic reduction of the FTE normal equations, fused with their assembly.
//
// Replaces, for level 0 only, lm_assemble + bcr_factor + bcr_update + bcr_backsub (lm.cu, bcr.cu) - i.e. the linear
// algebra inside the reference's `opt.solve(m)` (/root/reference/src/all_optimizations.py:503-524).
// At level 0 the couplings between neighbouring super-blocks are still the raw smoothness band (:369-391): 3 x 3 blocks
// of DIAGONAL 25 x 25 matrices with at most 6 non-zero diagonals.  So instead of the dense route
//     P = R^-1 Lc_e, Q = R^-1 Lc_c^T (two 75^3/2 triangular solves), then Q^T Q, P^T P, Q^T P (three 75^3 products)
// an eliminated block e only needs W = D_e^-1 (75 x 75, symmetric) and every Schur term becomes a sparse sandwich
//     (C W C'^T)[(a,p),(a',q)] = sum_{b,b'} c(a,b,p) W[(b,p),(b',q)] c'(a',b',q)        <= 9 terms per entry.
//   l0_invert   (CTA per eliminated block)   builds D_e straight from the fp32 frame blocks H_n + the band into REGISTER
//               tiles (thread (ty,tx) of a 15 x 15 grid owns a 5 x 5 tile), inverts it by blocked Gauss-Jordan (15 block
//               pivots; the 5 x 5 pivot block is inverted in registers by the thread that owns it right after its own
//               update, so the chain of dependent fp64 divisions overlaps the other tiles' rank-5 update), writes W and
//               z = W b_e.  No D / Lc / P / Q traffic for the eliminated half of the blocks.
//   l0_update   (CTA per surviving block)    assembles D_j, rhs_j and applies the two sandwiches, and writes the new
//               coupling Lc_j = -C_{j,el} W_el C_{el,j-2} to the previous survivor.
//   l0_backsub  (CTA per eliminated block)   x_e = z_e - W_e (C_{e,a} x_a + C_{c,e}^T x_c)   - a mat-vec, no triangular solve.
// 0.45 MFMA per eliminated block instead of 2.3 MFMA, and the dense levels below start from half the blocks.
#include "lm_common.cuh"

namespace acino {

constexpr int L0_THREADS = 256;
constexpr int TG = 15;            // 15 x 15 grid of 5 x 5 tiles
constexpr int TS = 5;

// in-place Gauss-Jordan inverse of an SPD 5 x 5 block held in registers; returns true on a non-positive pivot
__device__ __forceinline__ bool inv5(double A[TS][TS]) {
    bool bad = false;
#pragma unroll
    for (int q = 0; q < TS; ++q) {
        const double d = A[q][q];
        bad |= !(d > 0.0);
        const double inv = 1.0 / d;
#pragma unroll
        for (int j = 0; j < TS; ++j)
            if (j != q) A[q][j] *= inv;
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            if (i == q) continue;
            const double f = A[i][q];
#pragma unroll
            for (int j = 0; j < TS; ++j)
                if (j != q) A[i][j] = fma(-f, A[q][j], A[i][j]);
            A[i][q] = -f * inv;
        }
        A[q][q] = inv;
    }
    return bad;
}

__global__ void __launch_bounds__(L0_THREADS, 2)
l0_invert_kernel(const LmShard sh, const int* __restrict__ elim /*[ne][3]*/, const float* __restrict__ H,
                 const double* __restrict__ gtot, const unsigned char* __restrict__ fixed, const double* __restrict__ sw,
                 const double* __restrict__ ctl, double* __restrict__ W /* = P */, double* __restrict__ rhs,
                 int* __restrict__ info) {
    __shared__ double Gs[2][TS * TS];
    __shared__ double colp[SBN * TS];          // old A_IK: [row][m]
    __shared__ double rowp[TS * (SBN + 1)];    // new A_KJ = G A_KJ: [m][col]
    __shared__ double bvec[SBN];
    __shared__ double zp[TG][SBN + 1];
    const int e = elim[3 * blockIdx.x];
    const int tid = threadIdx.x;
    const bool tile = tid < TG * TG;
    const int ty = tid / TG, tx = tid - ty * TG;
    const double lambda = ctl[CTL_LAM];
    double A[TS][TS];
    if (tile) {
        const int a = ty / TS, p0 = TS * (ty - a * TS), b = tx / TS, q0 = TS * (tx - b * TS);
#pragma unroll
        for (int i = 0; i < TS; ++i)
#pragma unroll
            for (int j = 0; j < TS; ++j) A[i][j] = diag_entry(sh, H, fixed, sw, lambda, e, a, p0 + i, b, q0 + j);
    }
    if (tid < SBN) {
        const int n = 3 * e + tid / NA;
        const size_t i = (size_t)n * NA + tid % NA;
        bvec[tid] = (n < sh.n_frames && !fixed[i]) ? -gtot[i] : 0.0;
    }
    if (tile && ty == 0 && tx == 0) {
        if (inv5(A)) atomicExch(info, e + 1);
#pragma unroll
        for (int i = 0; i < TS; ++i)
#pragma unroll
            for (int j = 0; j < TS; ++j) Gs[0][i * TS + j] = A[i][j];
    }
    __syncthreads();
    for (int k = 0; k < TG; ++k) {
        // ---- panels of pivot block k
        if (tile && (ty == k) != (tx == k)) {
            const double* G = Gs[k & 1];
            if (ty == k) {          // row panel: A_KJ <- G A_KJ (column by column), published for the rank-5 update
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    double t[TS];
#pragma unroll
                    for (int i = 0; i < TS; ++i) {
                        double s = 0.0;
#pragma unroll
                        for (int m = 0; m < TS; ++m) s = fma(G[i * TS + m], A[m][j], s);
                        t[i] = s;
                    }
#pragma unroll
                    for (int i = 0; i < TS; ++i) {
                        A[i][j] = t[i];
                        rowp[i * (SBN + 1) + TS * tx + j] = t[i];
                    }
                }
            } else {                // column panel: publish the old A_IK, then A_IK <- -A_IK G (row by row)
#pragma unroll
                for (int i = 0; i < TS; ++i) {
                    double t[TS];
#pragma unroll
                    for (int j = 0; j < TS; ++j) {
                        colp[(TS * ty + i) * TS + j] = A[i][j];
                        double s = 0.0;
#pragma unroll
                        for (int m = 0; m < TS; ++m) s = fma(A[i][m], G[m * TS + j], s);
                        t[j] = -s;
                    }
#pragma unroll
                    for (int j = 0; j < TS; ++j) A[i][j] = t[j];
                }
            }
        }
        __syncthreads();
        // ---- rank-5 update of every other tile: A_IJ -= A_IK(old) (G A_KJ)
        if (tile && ty != k && tx != k) {
#pragma unroll
            for (int m = 0; m < TS; ++m) {
                double cm[TS], rm[TS];
#pragma unroll
                for (int i = 0; i < TS; ++i) cm[i] = colp[(TS * ty + i) * TS + m];
#pragma unroll
                for (int j = 0; j < TS; ++j) rm[j] = rowp[m * (SBN + 1) + TS * tx + j];
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) A[i][j] = fma(-cm[i], rm[j], A[i][j]);
            }
            if (ty == k + 1 && tx == k + 1) {      // next pivot block: invert it while the others finish their update
                if (inv5(A)) atomicExch(info, e + 1);
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) Gs[(k + 1) & 1][i * TS + j] = A[i][j];
            }
        }
        __syncthreads();
    }
    // ---- z = W b (fixed summation order), W written symmetric from its upper tiles
    if (tile) {
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < TS; ++j) s = fma(A[i][j], bvec[TS * tx + j], s);
            zp[tx][TS * ty + i] = s;
        }
        double* We = W + (size_t)e * SBN * SBN;
        if (ty <= tx) {
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    if (ty == tx && j < i) continue;
                    const int r = TS * ty + i, c = TS * tx + j;
                    We[r * SBN + c] = A[i][j];
                    We[c * SBN + r] = A[i][j];
                }
        }
    }
    __syncthreads();
    if (tid < SBN) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < TG; ++t) s += zp[t][tid];
        rhs[(size_t)e * SBN + tid] = s;
    }
}

// surviving block j with eliminated neighbours el = j - 1 / er = j + 1 (or -1)
__global__ void __launch_bounds__(L0_THREADS)
l0_update_kernel(const LmShard sh, const int* __restrict__ surv /*[ns][3]*/, const float* __restrict__ H,
                 const double* __restrict__ gtot, const unsigned char* __restrict__ fixed, const double* __restrict__ sw,
                 const double* __restrict__ ctl, const double* __restrict__ W /* = P */, double* __restrict__ D,
                 double* __restrict__ Lc, double* __restrict__ rhs) {
    __shared__ double cL[3][3][NA];     // coupling (block j frame a) - (block j-1 frame b)
    __shared__ double cR[3][3][NA];     // coupling (block j frame a) - (block j+1 frame b)
    __shared__ double cLL[3][3][NA];    // coupling (block j-1 frame b') - (block j-2 frame a')
    __shared__ double zl[SBN], zr[SBN];
    const int j = surv[3 * blockIdx.x], el = surv[3 * blockIdx.x + 1], er = surv[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    const double lambda = ctl[CTL_LAM];
    if (tid < 3 * 3 * NA) {
        const int a = tid / (3 * NA), b = (tid / NA) % 3, p = tid % NA;
        cL[a][b][p] = band_coupling(sh, fixed, sw, j, a, b, p);
        cR[a][b][p] = er >= 0 ? band_coupling(sh, fixed, sw, j + 1, b, a, p) : 0.0;
        cLL[a][b][p] = el >= 0 ? band_coupling(sh, fixed, sw, j - 1, a, b, p) : 0.0;
    }
    if (tid < SBN) {
        zl[tid] = el >= 0 ? rhs[(size_t)el * SBN + tid] : 0.0;
        zr[tid] = er >= 0 ? rhs[(size_t)er * SBN + tid] : 0.0;
    }
    __syncthreads();
    const double* Wl = W + (size_t)(el >= 0 ? el : 0) * SBN * SBN;
    const double* Wr = W + (size_t)(er >= 0 ? er : 0) * SBN * SBN;
    double* Dj = D + (size_t)j * SBN * SBN;
    double* Lj = Lc + (size_t)j * SBN * SBN;
    for (int t = tid; t < SBN * SBN; t += L0_THREADS) {
        const int r = t / SBN, c = t - r * SBN;
        const int a = r / NA, p = r - a * NA, a2 = c / NA, q = c - a2 * NA;
        // ---- coupling to the previous survivor
        double lv;
        if (el >= 0) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const double cl = cL[a][b][p];
                if (cl == 0.0) continue;
#pragma unroll
                for (int b2 = 0; b2 < 3; ++b2) {
                    const double cll = cLL[b2][a2][q];
                    if (cll == 0.0) continue;
                    acc = fma(cl * cll, Wl[(b * NA + p) * SBN + b2 * NA + q], acc);
                }
            }
            lv = -acc;
        } else {
            lv = (p == q) ? cL[a][a2][p] : 0.0;
        }
        Lj[t] = lv;
        // ---- diagonal block, upper triangle + mirror image (stays exactly symmetric)
        if (c >= r) {
            double v = diag_entry(sh, H, fixed, sw, lambda, j, a, p, a2, q);
            if (el >= 0) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const double c1 = cL[a][b][p];
                    if (c1 == 0.0) continue;
#pragma unroll
                    for (int b2 = 0; b2 < 3; ++b2) {
                        const double c2 = cL[a2][b2][q];
                        if (c2 == 0.0) continue;
                        acc = fma(c1 * c2, Wl[(b * NA + p) * SBN + b2 * NA + q], acc);
                    }
                }
                v -= acc;
            }
            if (er >= 0) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const double c1 = cR[a][b][p];
                    if (c1 == 0.0) continue;
#pragma unroll
                    for (int b2 = 0; b2 < 3; ++b2) {
                        const double c2 = cR[a2][b2][q];
                        if (c2 == 0.0) continue;
                        acc = fma(c1 * c2, Wr[(b * NA + p) * SBN + b2 * NA + q], acc);
                    }
                }
                v -= acc;
            }
            Dj[r * SBN + c] = v;
            Dj[c * SBN + r] = v;
        }
    }
    if (tid < SBN) {
        const int a = tid / NA, p = tid - a * NA;
        const int n = 3 * j + a;
        const size_t i = (size_t)n * NA + p;
        double v = (n < sh.n_frames && !fixed[i]) ? -gtot[i] : 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) v -= cL[a][b][p] * zl[b * NA + p] + cR[a][b][p] * zr[b * NA + p];   // zl / zr are 0 without el / er
        rhs[(size_t)j * SBN + tid] = v;
    }
}

// x_e = z_e - W_e (C_{e,a} x_a + C_{c,e}^T x_c) for the blocks eliminated at level 0
__global__ void __launch_bounds__(128)
l0_backsub_kernel(const LmShard sh, const int* __restrict__ elim, const unsigned char* __restrict__ fixed,
                  const double* __restrict__ sw, const double* __restrict__ W, const double* __restrict__ rhs,
                  double* __restrict__ x) {
    __shared__ double v[SBN];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    if (tid < SBN) {
        const int ea = tid / NA, p = tid - ea * NA;
        double s = 0.0;
        if (a >= 0)
#pragma unroll
            for (int b = 0; b < 3; ++b) s = fma(band_coupling(sh, fixed, sw, e, ea, b, p), x[(size_t)a * SBN + b * NA + p], s);
        if (c >= 0)
#pragma unroll
            for (int b = 0; b < 3; ++b) s = fma(band_coupling(sh, fixed, sw, e + 1, b, ea, p), x[(size_t)c * SBN + b * NA + p], s);
        v[tid] = s;
    }
    __syncthreads();
    if (tid < SBN) {
        const double* We = W + (size_t)e * SBN * SBN;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 5
        for (int k = 0; k < NA; ++k) {          // W is symmetric: column tid read as row-major W[k][tid] (coalesced)
            s0 = fma(We[k * SBN + tid], v[k], s0);
            s1 = fma(We[(k + NA) * SBN + tid], v[k + NA], s1);
            s2 = fma(We[(k + 2 * NA) * SBN + tid], v[k + 2 * NA], s2);
        }
        x[(size_t)e * SBN + tid] = rhs[(size_t)e * SBN + tid] - ((s0 + s1) + s2);
    }
}

cudaError_t launch_l0_invert(const LmShard& sh, int n_elim, const int* elim, const float* H, const double* gtot,
                             const unsigned char* fixed, const double* sw, const double* ctl, double* W, double* rhs,
                             int* info, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    l0_invert_kernel<<<n_elim, L0_THREADS, 0, s>>>(sh, elim, H, gtot, fixed, sw, ctl, W, rhs, info);
    return cudaGetLastError();
}

cudaError_t launch_l0_update(const LmShard& sh, int n_surv, const int* surv, const float* H, const double* gtot,
                             const unsigned char* fixed, const double* sw, const double* ctl, const double* W, double* D,
                             double* Lc, double* rhs, cudaStream_t s) {
    if (n_surv <= 0) return cudaSuccess;
    l0_update_kernel<<<n_surv, L0_THREADS, 0, s>>>(sh, surv, H, gtot, fixed, sw, ctl, W, D, Lc, rhs);
    return cudaGetLastError();
}

cudaError_t launch_l0_backsub(const LmShard& sh, int n_elim, const int* elim, const unsigned char* fixed, const double* sw,
                              const double* W, const double* rhs, double* x, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    l0_backsub_kernel<<<n_elim, 128, 0, s>>>(sh, elim, fixed, sw, W, rhs, x);
    return cudaGetLastError();
}

}  // namespace acino
