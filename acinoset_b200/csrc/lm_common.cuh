// Shared definitions of the FTE Levenberg-Marquardt kernels (lm.cu, lm_l0.cu, lm_plan.cu).
#pragma once
#include "acino_common.cuh"

namespace acino {

constexpr int SBF = 3;            // frames per super-block
constexpr int SBN = SBF * NA;     // 75 unknowns per super-block

// coefficient of D3^T D3 between global frames a and a+k (0 <= k <= 3), D3 rows m = 3..ng-1 with
// stencil (-1, 3, -3, 1) on columns m-3..m  (backwards_euler_pos / _vel + constant_acc eliminated,
// /root/reference/src/all_optimizations.py:369-391)
__host__ __device__ __forceinline__ double d3_stencil(long long i) {   // (-1, 3, -3, 1), no local-memory table
    return i == 0 ? -1.0 : (i == 1 ? 3.0 : (i == 2 ? -3.0 : 1.0));
}
__host__ __device__ __forceinline__ double band_coef(long long a, int k, long long ng) {
    if (a < 0 || a + k >= ng) return 0.0;
    long long m0 = a + k;
    if (m0 < 3) m0 = 3;
    long long m1 = a + 3;
    if (m1 > ng - 1) m1 = ng - 1;
    double s = 0.0;
    for (long long m = m0; m <= m1; ++m) s += d3_stencil(a - m + 3) * d3_stencil(a + k - m + 3);
    return s;
}

// Shard of the trajectory a rank works on (frame indices in band_coef are GLOBAL).
struct LmShard {
    int n_frames;            // local frames N
    long long frame0, ng;    // first global frame of the shard, global frame count
};

// Entry [(a,p),(b,p)] of the coupling block between super-block X (rows, local frame 3X + a) and super-block
// X - 1 (columns, local frame 3X - 3 + b): the smoothness band, zero for frozen variables, padding frames and
// beyond the global start.  Identical to what lm_assemble_kernel writes into Lc[X].  `fixed` lookups outside the
// shard (the neighbouring rank's frames) count as free: the interface chain masks them after the exchange.
__device__ __forceinline__ double band_coupling(const LmShard& sh, const unsigned char* __restrict__ fixed,
                                                const double* __restrict__ sw, const int X, const int a, const int b,
                                                const int p) {
    const int kk = 3 + a - b;
    if (kk > 3) return 0.0;
    const int nx = 3 * X + a, ny = 3 * X - 3 + b;
    if (nx >= sh.n_frames) return 0.0;
    if (sh.frame0 + ny < 0 || (X == 0 && sh.frame0 == 0)) return 0.0;
    if (fixed[(size_t)nx * NA + p]) return 0.0;
    if (ny >= 0 && ny < sh.n_frames && fixed[(size_t)ny * NA + p]) return 0.0;
    return band_coef(sh.frame0 + ny, kk, sh.ng) * sw[p];
}

// Entry [(a,p),(b,q)] of the diagonal super-block X of B + lam diag(B) (what lm_assemble_kernel writes into D[X]).
__device__ __forceinline__ double diag_entry(const LmShard& sh, const float* __restrict__ H,
                                             const unsigned char* __restrict__ fixed, const double* __restrict__ sw,
                                             const double lambda, const int X, const int a, const int p, const int b,
                                             const int q) {
    const int na = 3 * X + a, nb = 3 * X + b;
    const bool same = (a == b) && (p == q);
    if (na >= sh.n_frames || nb >= sh.n_frames) return same ? 1.0 : 0.0;            // padding frame
    if (fixed[(size_t)na * NA + p] || fixed[(size_t)nb * NA + q]) return same ? 1.0 : 0.0;   // frozen variable
    if (a == b) {
        const int lo_ = p < q ? p : q, hi_ = p < q ? q : p;
        const double h = (double)H[(size_t)na * NU + upper_index(lo_, hi_)];
        if (p != q) return h;
        return (h + band_coef(sh.frame0 + na, 0, sh.ng) * sw[p]) * (1.0 + lambda);  // Marquardt: B_pp + lam B_pp
    }
    if (p != q) return 0.0;
    const int k = a > b ? a - b : b - a;
    return band_coef(sh.frame0 + (a < b ? na : nb), k, sh.ng) * sw[p];
}

// Device-resident control block of one LM solve (doubles; integers are stored as doubles).
enum LmCtl {
    CTL_LAM = 0, CTL_F, CTL_FT, CTL_PRED, CTL_STEP, CTL_RHO, CTL_REL, CTL_ACCEPT, CTL_DONE, CTL_N_ATTEMPT, CTL_N_ACCEPT,
    CTL_FAIL_STREAK, CTL_STATUS, CTL_MAX_ITER, CTL_MAX_ATTEMPTS, CTL_TOL_STEP, CTL_TOL_REL, CTL_ITERS, CTL_HIST_CAP,
    CTL_TOL_NOISE, CTL_NOISE_STREAK, CTL_N_ENQ, CTL_DONE_AT,
    CTL_SIZE = 32
};
constexpr int LM_SUMS = 8;        // per-rank partial sums exchanged per attempt: cost, cost_s, pred, -, max step, ...
constexpr int LM_HIST = 8;        // per-attempt log record: F, Ft, lam, rho, step, accepted, pred, -

}  // namespace acino
