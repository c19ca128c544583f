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

// ---- inputs of one super-block staged in shared memory ----------------------------------------------------------------
// The global-memory versions (diag_entry / band_coupling in lm_common.cuh) chain two dependent L2 round trips per matrix
// entry (frozen flag -> branch -> H entry); 25 entries per thread made the BUILD a third of l0_invert.  Here every input
// is fetched once, coalesced, and the entries are formed from shared memory.
struct BandTab {
    double diag[3][3];            // band_coef between frames a, b of the block itself (|a - b| <= 2)
    double left[3][3];            // band_coef between frame a of block X and frame b of block X - 1 (0 if more than 3 apart,
                                  // beyond the global start or a padding frame)
};
__device__ __forceinline__ void band_tab(const LmShard& sh, const int X, BandTab& t, const int idx /*0..17*/) {
    const int which = idx / 9, a = (idx % 9) / 3, b = idx % 3;
    if (which == 0) {
        const int na = 3 * X + a, nb = 3 * X + b;
        const int k = a > b ? a - b : b - a;
        t.diag[a][b] = (na < sh.n_frames && nb < sh.n_frames) ? band_coef(sh.frame0 + (a < b ? na : nb), k, sh.ng) : 0.0;
    } else {
        const int kk = 3 + a - b, nx = 3 * X + a, ny = 3 * X - 3 + b;
        const bool ok = kk <= 3 && X >= 0 && nx < sh.n_frames && sh.frame0 + ny >= 0 && !(X == 0 && sh.frame0 == 0);
        t.left[a][b] = ok ? band_coef(sh.frame0 + ny, kk, sh.ng) : 0.0;
    }
}
// frozen flags of super-block X (0 outside the shard: the neighbouring rank's frames count as free, lm_common.cuh)
__device__ __forceinline__ unsigned char fixed_at(const LmShard& sh, const unsigned char* __restrict__ fixed, const int X,
                                                  const int t /*0..74*/) {
    const int n = 3 * X + t / NA;
    return (X >= 0 && n < sh.n_frames) ? fixed[(size_t)n * NA + t % NA] : (unsigned char)0;
}

// in-place Gauss-Jordan inverse of an SPD 5 x 5 block held in registers; returns true on a non-positive pivot.
// (IEEE division: 80 cycles dependent on B200; a MUFU seed + two Newton steps measured 92 - scripts/micro/l0_bench.cu.)
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

#ifdef ACINO_L0_TIMING
__device__ long long g_l0_cycles[8];
#define L0_MARK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _t = clock64(); atomicAdd((unsigned long long*)&g_l0_cycles[slot], (unsigned long long)(_t - _tprev)); _tprev = _t; } } while (0)
#else
#define L0_MARK(slot) do { } while (0)
#endif

// entry [(a,p),(b,q)] of the diagonal super-block from staged inputs: Hs [3][NU] fp32 frame blocks, fx [75] frozen flags
__device__ __forceinline__ double diag_entry_s(const float* __restrict__ Hs, const unsigned char* __restrict__ fx,
                                               const double* __restrict__ swv, const BandTab& bt, const int n_valid,
                                               const double lambda, const int a, const int p, const int b, const int q) {
    const bool same = (a == b) && (p == q);
    if (a >= n_valid || b >= n_valid || fx[a * NA + p] || fx[b * NA + q]) return same ? 1.0 : 0.0;
    if (a == b) {
        const int lo_ = p < q ? p : q, hi_ = p < q ? q : p;
        const double h = (double)Hs[a * NU + upper_index(lo_, hi_)];
        return p != q ? h : (h + bt.diag[a][a] * swv[p]) * (1.0 + lambda);
    }
    return p == q ? bt.diag[a][b] * swv[p] : 0.0;
}

constexpr size_t L0_INVERT_SMEM = (size_t)SBN * SBN * sizeof(double);

__global__ void __launch_bounds__(L0_THREADS, 2)
l0_invert_kernel(const LmShard sh, const int* __restrict__ elim /*[ne][3]*/, const float* __restrict__ H,
                 const double* __restrict__ gtot, const unsigned char* __restrict__ fixed, const double* __restrict__ sw,
                 const double* __restrict__ ctl, double* __restrict__ W /* = P */, double* __restrict__ rhs,
                 int* __restrict__ info) {
    __shared__ double Gs[2][TS * TS];
    __shared__ double colp[2][SBN * TS];       // old A_IK:             [row][m]   (double buffered: written for pivot k+1
                                               //                                   while the update of pivot k reads it)
    __shared__ double cnew[SBN * TS];          // new A_IK = -A_IK G:   [row][m]
    __shared__ double rraw[TS * (SBN + 1)];    // old A_KJ:             [m][col]
    __shared__ double rowp[TS * (SBN + 1)];    // new A_KJ = G A_KJ:    [m][col]
    __shared__ double bvec[SBN];
    __shared__ double zp[TG][SBN + 1];         // also the staging area of the inputs (3 frame blocks of H) before the loop
    __shared__ double swv[NA];
    __shared__ BandTab bt;
    __shared__ unsigned char fx[SBN + 1];
    extern __shared__ __align__(16) double wst[];      // [75][75] W staged for a coalesced, exactly symmetric write-out
    const int e = elim[3 * blockIdx.x];
    const int tid = threadIdx.x;
    const bool tile = tid < TG * TG;
    const int ty = tid / TG, tx = tid - ty * TG;
    const double lambda = ctl[CTL_LAM];
#ifdef ACINO_L0_TIMING
    long long _tprev = clock64();
#endif
    // ---- stage the block's inputs (coalesced), then build the register tiles from shared memory
    float* Hs = reinterpret_cast<float*>(&zp[0][0]);          // 3 * 325 floats = 3.9 KB of the 9 KB zp
    const int n_valid = min(3, sh.n_frames - 3 * e);
    for (int i = tid; i < n_valid * NU; i += L0_THREADS) Hs[i] = H[(size_t)3 * e * NU + i];
    if (tid < SBN) {
        const unsigned char f = fixed_at(sh, fixed, e, tid);
        fx[tid] = f;
        const int n = 3 * e + tid / NA;
        bvec[tid] = (n < sh.n_frames && !f) ? -gtot[(size_t)n * NA + tid % NA] : 0.0;
    } else if (tid >= 96 && tid < 96 + NA) {
        swv[tid - 96] = sw[tid - 96];
    } else if (tid >= 128 && tid < 128 + 9) {
        band_tab(sh, e, bt, tid - 128);
    }
    __syncthreads();
    double A[TS][TS];
    if (tile) {
        const int a = ty / TS, p0 = TS * (ty - a * TS), b = tx / TS, q0 = TS * (tx - b * TS);
#pragma unroll
        for (int i = 0; i < TS; ++i)
#pragma unroll
            for (int j = 0; j < TS; ++j) A[i][j] = diag_entry_s(Hs, fx, swv, bt, n_valid, lambda, a, p0 + i, b, q0 + j);
        // pivot block 0 is inverted by its owner; the owners of its panels publish them
        if (ty == 0 && tx == 0) {
            if (inv5(A)) atomicExch(info, e + 1);
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) Gs[0][i * TS + j] = A[i][j];
        } else if (ty == 0) {
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) rraw[i * (SBN + 1) + TS * tx + j] = A[i][j];
        } else if (tx == 0) {
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) colp[0][(TS * ty + i) * TS + j] = A[i][j];
        }
    }
    __syncthreads();
    L0_MARK(0);
    for (int k = 0; k < TG; ++k) {
        const double* cp = colp[k & 1];
        // ---- (1) cooperative panel products: threads 0..74 one COLUMN of G A_K,: ; threads 128..202 one ROW of -A_:,K G
        {
            const double* G = Gs[k & 1];
            if (tid < SBN) {
                double v[TS];
#pragma unroll
                for (int m = 0; m < TS; ++m) v[m] = rraw[m * (SBN + 1) + tid];
#pragma unroll
                for (int i = 0; i < TS; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m < TS; ++m) s = fma(G[i * TS + m], v[m], s);
                    rowp[i * (SBN + 1) + tid] = s;
                }
            } else if (tid >= 128 && tid < 128 + SBN) {
                const int r = tid - 128;
                double v[TS];
#pragma unroll
                for (int m = 0; m < TS; ++m) v[m] = cp[r * TS + m];
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m < TS; ++m) s = fma(v[m], G[m * TS + j], s);
                    cnew[r * TS + j] = -s;
                }
            }
        }
        __syncthreads();
        L0_MARK(1);
        // ---- (2) rank-5 update of every other tile: A_IJ -= A_IK(old) (G A_KJ); the owners of pivot block k's panels pick
        //      up their new tiles; the owners of pivot block k+1 (tile, row panel, column panel) invert / publish theirs
        if (tile) {
            if (ty != k && tx != k) {
#pragma unroll
                for (int m = 0; m < TS; ++m) {
                    double cm[TS], rm[TS];
#pragma unroll
                    for (int i = 0; i < TS; ++i) cm[i] = cp[(TS * ty + i) * TS + m];
#pragma unroll
                    for (int j = 0; j < TS; ++j) rm[j] = rowp[m * (SBN + 1) + TS * tx + j];
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) A[i][j] = fma(-cm[i], rm[j], A[i][j]);
                }
            } else if (ty == k && tx != k) {
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) A[i][j] = rowp[i * (SBN + 1) + TS * tx + j];
            } else if (tx == k && ty != k) {
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) A[i][j] = cnew[(TS * ty + i) * TS + j];
            }
            const int k1 = k + 1;
            if (k1 < TG) {
                if (ty == k1 && tx == k1) {
                    if (inv5(A)) atomicExch(info, e + 1);
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) Gs[k1 & 1][i * TS + j] = A[i][j];
                } else if (ty == k1) {
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) rraw[i * (SBN + 1) + TS * tx + j] = A[i][j];
                } else if (tx == k1) {
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) colp[k1 & 1][(TS * ty + i) * TS + j] = A[i][j];
                }
            }
        }
        __syncthreads();
        L0_MARK(2);
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
        if (ty <= tx) {       // upper tiles define both triangles
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    if (ty == tx && j < i) continue;
                    const int r = TS * ty + i, c = TS * tx + j;
                    wst[r * SBN + c] = A[i][j];
                    wst[c * SBN + r] = A[i][j];
                }
        }
    }
    __syncthreads();
    {
        double* We = W + (size_t)e * SBN * SBN;
        for (int i = tid; i < SBN * SBN; i += L0_THREADS) We[i] = wst[i];
    }
    if (tid < SBN) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < TG; ++t) s += zp[t][tid];
        rhs[(size_t)e * SBN + tid] = s;
    }
    L0_MARK(3);
}

#ifdef ACINO_L0_TIMING
extern "C" void acino_debug_l0_cycles(long long* out8) { cudaMemcpyFromSymbol(out8, g_l0_cycles, sizeof(long long) * 8); }
extern "C" void acino_debug_l0_reset() { long long z[8] = {0}; cudaMemcpyToSymbol(g_l0_cycles, z, sizeof(z)); }
#endif

// surviving block j with eliminated neighbours el = j - 1 / er = j + 1 (or -1).  W_el and W_er are staged in shared
// memory (8-byte async copies: a block of P starts on an 8-byte boundary only); the 3 x 3 frame-block structure is
// walked with uniform loop bounds - c(a,b,.) of a left coupling is zero for b < a, of a right coupling for b > a -
// and thread <-> (p,q) inside a frame block, q fastest: conflict-free shared-memory reads, coalesced stores.
constexpr size_t L0_UPDATE_SMEM = (size_t)(2 * SBN * SBN) * sizeof(double);

__device__ __forceinline__ void cp_async8_l0(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}

__global__ void __launch_bounds__(L0_THREADS, 2)
l0_update_kernel(const LmShard sh, const int* __restrict__ surv /*[ns][3]*/, const float* __restrict__ H,
                 const double* __restrict__ gtot, const unsigned char* __restrict__ fixed, const double* __restrict__ sw,
                 const double* __restrict__ ctl, const double* __restrict__ W /* = P */, double* __restrict__ D,
                 double* __restrict__ Lc, double* __restrict__ rhs) {
    extern __shared__ __align__(16) double smw[];
    double* Wl = smw;                   // W_el
    double* Wr = smw + SBN * SBN;       // W_er
    __shared__ double cL[3][3][NA];     // coupling (block j frame a) - (block j-1 frame b)
    __shared__ double cR[3][3][NA];     // coupling (block j frame a) - (block j+1 frame b)
    __shared__ double cLL[3][3][NA];    // coupling (block j-1 frame b') - (block j-2 frame a')
    __shared__ double zl[SBN], zr[SBN], gj[SBN], swv[NA];
    __shared__ float Hs[3 * NU];
    __shared__ BandTab bt[3];           // [0] block j-1, [1] block j, [2] block j+1
    __shared__ unsigned char fx[4][SBN + 1];   // frozen flags of blocks j-2 .. j+1
    const int j = surv[3 * blockIdx.x], el = surv[3 * blockIdx.x + 1], er = surv[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    const double lambda = ctl[CTL_LAM];
    const int n_valid = min(3, sh.n_frames - 3 * j);
    {
        const double* gl = W + (size_t)(el >= 0 ? el : 0) * SBN * SBN;
        const double* gr = W + (size_t)(er >= 0 ? er : 0) * SBN * SBN;
        for (int i = tid; i < SBN * SBN; i += L0_THREADS) {
            if (el >= 0) cp_async8_l0(&Wl[i], gl + i);
            if (er >= 0) cp_async8_l0(&Wr[i], gr + i);
        }
    }
    for (int i = tid; i < n_valid * NU; i += L0_THREADS) Hs[i] = H[(size_t)3 * j * NU + i];
    for (int i = tid; i < 4 * SBN; i += L0_THREADS) fx[i / SBN][i % SBN] = fixed_at(sh, fixed, j - 2 + i / SBN, i % SBN);
    if (tid < SBN) {
        zl[tid] = el >= 0 ? rhs[(size_t)el * SBN + tid] : 0.0;
        zr[tid] = er >= 0 ? rhs[(size_t)er * SBN + tid] : 0.0;
        const int n = 3 * j + tid / NA;
        gj[tid] = n < sh.n_frames ? gtot[(size_t)n * NA + tid % NA] : 0.0;
    } else if (tid >= 96 && tid < 96 + NA) {
        swv[tid - 96] = sw[tid - 96];
    } else if (tid >= 128 && tid < 128 + 54) {
        const int which = (tid - 128) / 18;
        band_tab(sh, j - 1 + which, bt[which], (tid - 128) % 18);
    }
    __syncthreads();
    if (tid < 3 * 3 * NA) {
        const int a = tid / (3 * NA), b = (tid / NA) % 3, p = tid % NA;
        // rows block X frame a, columns block X - 1 frame b: free on both sides (the band table holds the geometry)
        cL[a][b][p] = (fx[2][a * NA + p] || fx[1][b * NA + p]) ? 0.0 : bt[1].left[a][b] * swv[p];
        cR[a][b][p] = (er < 0 || fx[3][b * NA + p] || fx[2][a * NA + p]) ? 0.0 : bt[2].left[b][a] * swv[p];
        cLL[a][b][p] = (el < 0 || fx[1][a * NA + p] || fx[0][b * NA + p]) ? 0.0 : bt[0].left[a][b] * swv[p];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    double* Dj = D + (size_t)j * SBN * SBN;
    double* Lj = Lc + (size_t)j * SBN * SBN;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int a2 = 0; a2 < 3; ++a2) {
            for (int t = tid; t < NA * NA; t += L0_THREADS) {
                const int p = t / NA, q = t - p * NA;
                const int r = a * NA + p, c = a2 * NA + q;
                // ---- coupling to the previous survivor: -sum_{b >= a, b2 <= a2} cL[a][b][p] W_el[(b,p),(b2,q)] cLL[b2][a2][q]
                double lv;
                if (el >= 0) {
                    double acc = 0.0;
#pragma unroll
                    for (int b = a; b < 3; ++b)
#pragma unroll
                        for (int b2 = 0; b2 <= a2; ++b2)
                            acc = fma(cL[a][b][p] * cLL[b2][a2][q], Wl[(b * NA + p) * SBN + b2 * NA + q], acc);
                    lv = -acc;
                } else {
                    lv = (p == q) ? cL[a][a2][p] : 0.0;
                }
                Lj[r * SBN + c] = lv;
                // ---- diagonal block, upper triangle + mirror image (stays exactly symmetric)
                if (a2 > a || (a2 == a && q >= p)) {
                    double v = diag_entry_s(Hs, fx[2], swv, bt[1], n_valid, lambda, a, p, a2, q);
                    if (el >= 0) {
                        double acc = 0.0;
#pragma unroll
                        for (int b = a; b < 3; ++b)
#pragma unroll
                            for (int b2 = a2; b2 < 3; ++b2)
                                acc = fma(cL[a][b][p] * cL[a2][b2][q], Wl[(b * NA + p) * SBN + b2 * NA + q], acc);
                        v -= acc;
                    }
                    if (er >= 0) {
                        double acc = 0.0;
#pragma unroll
                        for (int b = 0; b <= a; ++b)
#pragma unroll
                            for (int b2 = 0; b2 <= a2; ++b2)
                                acc = fma(cR[a][b][p] * cR[a2][b2][q], Wr[(b * NA + p) * SBN + b2 * NA + q], acc);
                        v -= acc;
                    }
                    Dj[r * SBN + c] = v;
                    Dj[c * SBN + r] = v;
                }
            }
        }
    }
    if (tid < SBN) {
        const int a = tid / NA, p = tid - a * NA;
        double v = (a < n_valid && !fx[2][tid]) ? -gj[tid] : 0.0;
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
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {        // the attribute is per device
        cudaError_t e = cudaFuncSetAttribute(l0_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L0_INVERT_SMEM);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    l0_invert_kernel<<<n_elim, L0_THREADS, L0_INVERT_SMEM, s>>>(sh, elim, H, gtot, fixed, sw, ctl, W, rhs, info);
    return cudaGetLastError();
}

cudaError_t launch_l0_update(const LmShard& sh, int n_surv, const int* surv, const float* H, const double* gtot,
                             const unsigned char* fixed, const double* sw, const double* ctl, const double* W, double* D,
                             double* Lc, double* rhs, cudaStream_t s) {
    if (n_surv <= 0) return cudaSuccess;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {        // the attribute is per device
        cudaError_t e = cudaFuncSetAttribute(l0_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L0_UPDATE_SMEM);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    l0_update_kernel<<<n_surv, L0_THREADS, L0_UPDATE_SMEM, s>>>(sh, surv, H, gtot, fixed, sw, ctl, W, D, Lc, rhs);
    return cudaGetLastError();
}

cudaError_t launch_l0_backsub(const LmShard& sh, int n_elim, const int* elim, const unsigned char* fixed, const double* sw,
                              const double* W, const double* rhs, double* x, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    l0_backsub_kernel<<<n_elim, 128, 0, s>>>(sh, elim, fixed, sw, W, rhs, x);
    return cudaGetLastError();
}

}  // namespace acino
