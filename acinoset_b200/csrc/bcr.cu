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
constexpr size_t SB2 = (size_t)SB * SB;

// ---- bcr_factor: fused, blocked elimination of one super-block --------------------------------------
// Gaussian elimination (no pivoting; the block is SPD) of the augmented panel
//     [ D_e | Lc_e | Lc_c^T | b_e ]      75 x (75 + 75 + 75 + 1)
// held COLUMN-major in shared memory (column stride FCS = 76 doubles: 16-byte accesses of a quarter-warp
// = 4 columns x 2 row pairs are bank-conflict-free, pivot-column reads are broadcasts).  After eliminating
// pivots 0..k-1, row k of the right-hand columns is (L^-1 [Lc_e | Lc_c^T | b])_k with D = L Delta L^T, so
//     P = R^-1 Lc_e, Q = R^-1 Lc_c^T, z = R^-1 b   with R = L Delta^1/2   are row scalings by a_kk^-1/2:
// no explicit inverse and no separate GEMMs (0.5 MFMA per block instead of 1.1).
// A rank-1 update per pivot streams the whole trailing panel through shared memory 75 times and is bound
// by its 128 B/clk; so pivots are eliminated FB at a time (rank-FB update, the multipliers in registers):
//   (a) one thread factors the FB x FB diagonal block in registers, publishes 1/a_qq and c_pq = a_pq / a_qq
//       - for the NEXT block while (c) of the current one runs (look-ahead: the chain of dependent fp64
//       divisions is off the critical path)
//   (b) thread per row below the block: its FB panel entries (forward substitution with c);
//       thread per right-hand column: its FB pivot-row entries (same recurrence)
//   (c) two threads per trailing column (interleaved row pairs): a_ij += sum_p a_ip m_p, m_p = -a_pj / a_pp
//       (D columns: lower triangle only, a_pj read as a_jp)
// Stored for back-substitution in place of D_e: strict lower triangle = L (unit diagonal implied),
// diagonal = a_kk^-1/2.
constexpr int FCOLS = 3 * SB + 1;   // 226 panel columns
constexpr int FCS = 76;             // column stride (doubles): 152 words = 24 (mod 32)
constexpr int FB = 5;               // pivots per block step (divides 75)
constexpr int FACTOR_THREADS = 512;
constexpr size_t BCR_FACTOR_SMEM = (size_t)FCOLS * FCS * sizeof(double);
// Two-pass variant for the large levels (more blocks than SMs): pass 1 eliminates [D | Lc_e | b] (151 columns) and
// keeps the multipliers; pass 2 reloads the right-hand buffer with Lc_c^T and replays the elimination on those 75
// columns - every column is independent once the multipliers are known, so pass 2 needs no CTA barrier at all.
// 92 KB of shared memory instead of 137 KB: two CTAs per SM, which overlaps the latency chains of two blocks.
constexpr int FCOLS_2P = 2 * SB + 1;
constexpr size_t BCR_FACTOR2_SMEM = (size_t)FCOLS_2P * FCS * sizeof(double);

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}

#ifdef ACINO_BCR_TIMING
__device__ long long g_bcr_cycles[8];
#define BCR_MARK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _t = clock64(); atomicAdd((unsigned long long*)&g_bcr_cycles[slot], (unsigned long long)(_t - _tprev)); _tprev = _t; } } while (0)
#else
#define BCR_MARK(slot) do { } while (0)
#endif

// (a) of a diagonal block in registers, by ONE thread.  For k1 > 0 it first applies the pending rank-FB update of
// block k1 - FB to the FB x FB entries itself ("look-ahead"): the factorisation of the next diagonal block - a chain
// of 5 dependent fp64 divisions - then overlaps the trailing update (c) of the current block instead of serialising
// with it.  Publishes 1/a_qq (sinv) and c_pq = a_pq / a_qq (scpq, FB x FB row-major).
__device__ __forceinline__ void bcr_diag_block(double* __restrict__ sm, double* __restrict__ sinv, double* __restrict__ scpq,
                                               int* __restrict__ info, const int e, const int k1) {
    // every loop runs over the full 0..FB-1 range with a predicate: constant trip counts, so the 5 x 5 arrays
    // stay in registers (triangular loop bounds made the compiler index them in local memory)
    double A[FB][FB];
#pragma unroll
    for (int r = 0; r < FB; ++r)
#pragma unroll
        for (int p = 0; p < FB; ++p) A[r][p] = p <= r ? sm[(k1 + p) * FCS + k1 + r] : 0.0;
    if (k1 > 0) {
        const int k0 = k1 - FB;
        double l[FB][FB];          // l[q][r] = a_(k1+r)(k0+q): panel column q, row k1 + r
#pragma unroll
        for (int q = 0; q < FB; ++q)
#pragma unroll
            for (int r = 0; r < FB; ++r) l[q][r] = sm[(k0 + q) * FCS + k1 + r];
#pragma unroll
        for (int p = 0; p < FB; ++p) {
#pragma unroll
            for (int q = 0; q < FB; ++q) {
                const double mq = -(l[q][p] * sinv[k0 + q]);
#pragma unroll
                for (int r = 0; r < FB; ++r)
                    if (r >= p) A[r][p] = fma(l[q][r], mq, A[r][p]);
            }
        }
    }
    bool bad = false;
#pragma unroll
    for (int q = 0; q < FB; ++q) {
        const double d = A[q][q];
        bad |= !(d > 0.0);
        const double inv = 1.0 / d;
        sinv[k1 + q] = inv;
#pragma unroll
        for (int r = 0; r < FB; ++r) {
            if (r > q) {
                const double cq = A[r][q] * inv;          // a_rq / a_qq
                scpq[r * FB + q] = cq;
#pragma unroll
                for (int p = 0; p < FB; ++p)
                    if (p > q && p <= r) A[r][p] = fma(-cq, A[p][q], A[r][p]);
            }
        }
    }
    if (bad) atomicExch(info, e + 1);
#pragma unroll
    for (int r = 0; r < FB; ++r)
#pragma unroll
        for (int p = 0; p < FB; ++p)
            if (p <= r) sm[(k1 + p) * FCS + k1 + r] = A[r][p];
}

template <bool TWO_PASS>
__global__ void __launch_bounds__(FACTOR_THREADS, TWO_PASS ? 2 : 1)
bcr_factor_kernel(const int* __restrict__ elim /*[ne][3]*/, const double* __restrict__ D, const double* __restrict__ Lc,
                  double* __restrict__ P, double* __restrict__ Q, double* __restrict__ R, double* __restrict__ rhs,
                  int* __restrict__ info, const int split) {
    constexpr int NCOLS = TWO_PASS ? FCOLS_2P : FCOLS;      // panel columns of the (first) pass
    constexpr int ZCOL = NCOLS - 1;                          // the right-hand-side vector b is the last column
    extern __shared__ __align__(16) double sm[];             // [NCOLS][FCS]
    __shared__ double sinv[SB], sdi[SB], scpq_all[TWO_PASS ? SB / FB : 1][FB][FB];
    const int e = elim[3 * blockIdx.x], c = elim[3 * blockIdx.x + 2];
    // Split mode (small levels, two CTAs per block, TWO_PASS layout): slice 0 eliminates [D | Lc_e | b] and writes R, P, z;
    // slice 1 eliminates [D | Lc_c^T] and writes Q.  Both factor D redundantly - the SMs are idle at these levels - and each
    // streams a 151-column panel instead of 226 columns through shared memory (the kernel's bound).  The factor goes to R,
    // not back into D: the other slice may still be reading D_e.
    const int slice = (TWO_PASS && split) ? (int)blockIdx.y : 0;
    if (slice == 1 && c < 0) return;
    // neighbour whose coupling fills columns 75..149: the left one (Lc_e), or in slice 1 the right one (Lc_c^T)
    const int a = slice == 1 ? c : elim[3 * blockIdx.x + 1];
    const int tid = threadIdx.x;
#ifdef ACINO_BCR_TIMING
    long long _tprev = clock64();
#endif
    // ---- load the panel (column-major) with 8-byte async copies, all in flight at once
    {
        const double* gD = D + (size_t)e * SB2;
        const double* gA = Lc + (size_t)(slice == 1 ? c : e) * SB2;
        const double* gC = Lc + (size_t)(c >= 0 ? c : 0) * SB2;
        for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
            const int r = i / SB, cc = i - r * SB;
            cp_async8(&sm[r * FCS + cc], gD + i);                              // symmetric: (r,cc) == (cc,r)
            if (a >= 0) {
                if (slice == 1) cp_async8(&sm[(SB + r) * FCS + cc], gA + i);   // Lc_c^T (cc, r) = Lc_c (r, cc)
                else cp_async8(&sm[(SB + cc) * FCS + r], gA + i);              // Lc_e (r, cc)
            }
            if (!TWO_PASS && c >= 0) cp_async8(&sm[(2 * SB + r) * FCS + cc], gC + i);   // Lc_c^T (cc, r) = Lc_c (r, cc)
        }
        if (tid < SB) sm[ZCOL * FCS + tid] = slice == 1 ? 0.0 : rhs[(size_t)e * SB + tid];
        if (tid < NCOLS) sm[tid * FCS + SB] = 0.0;                             // row 75: zero padding
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    const int col = tid >> 1, half = tid & 1;
    const bool is_d = col < SB;
    const bool z_on = slice == 0;
    const bool rhs_on = (tid >= SB && tid < NCOLS) && (tid == ZCOL ? z_on : (tid < 2 * SB ? a >= 0 : c >= 0));     // phase (b)
    const bool active_col = col < NCOLS && (is_d || (col == ZCOL ? z_on : (col < 2 * SB ? a >= 0 : c >= 0)));       // phase (c)
    double* cj = sm + col * FCS;
    __syncthreads();
    // (a) by thread FACTOR_THREADS - 1 (its column slot is beyond the panel, so it is otherwise idle)
    if (tid == FACTOR_THREADS - 1) bcr_diag_block(sm, sinv, &scpq_all[0][0][0], info, e, 0);
    for (int k0 = 0; k0 < SB; k0 += FB) {
        __syncthreads();
        BCR_MARK(k0 == 0 ? 1 : 3);
        // ---- (b) panel rows below the block (threads 0..74) and pivot rows of the right-hand columns (threads 75..225)
        if (tid < SB ? tid >= k0 + FB : rhs_on) {
            // element p of this thread's vector: row `tid` of panel column k0+p, or row k0+p of column `tid`
            double* base = tid < SB ? sm + k0 * FCS + tid : sm + tid * FCS + k0;
            const int stride = tid < SB ? FCS : 1;
            double v[FB];
#pragma unroll
            for (int p = 0; p < FB; ++p) v[p] = base[p * stride];
#pragma unroll
            for (int p = 1; p < FB; ++p) {
#pragma unroll
                for (int q = 0; q < p; ++q) v[p] = fma(-scpq_all[TWO_PASS ? k0 / FB : 0][p][q], v[q], v[p]);
                base[p * stride] = v[p];
            }
        }
        __syncthreads();
        BCR_MARK(2);
        // ---- (c) rank-FB update of the trailing columns; meanwhile (a) of the next diagonal block
        const int k1 = k0 + FB;
        if (tid == FACTOR_THREADS - 1 && k1 < SB) bcr_diag_block(sm, sinv, &scpq_all[TWO_PASS ? k1 / FB : 0][0][0], info, e, k1);
        if (!active_col || col < k1) continue;
        double m[FB];
#pragma unroll
        for (int p = 0; p < FB; ++p) m[p] = -((is_d ? sm[(k0 + p) * FCS + col] : cj[k0 + p]) * sinv[k0 + p]);
        const double* pk = sm + k0 * FCS;
        // first row this column updates; the next diagonal block belongs to the look-ahead thread
        int i = is_d ? (col < k1 + FB ? k1 + FB : col) : k1;
        if (i & 1) {                             // odd first row: a single-row update, then aligned pairs
            if (half == 0 && i < SB) {
                double v = cj[i];
#pragma unroll
                for (int p = 0; p < FB; ++p) v = fma(pk[p * FCS + i], m[p], v);
                cj[i] = v;
            }
            ++i;
        }
        // aligned row pairs (i, i+1), interleaved over the column's two threads; the last pair touches padding
        i += 2 * half;
        for (; i + 4 < SB + 1; i += 8) {         // two pairs per trip
            double2 q0 = *reinterpret_cast<double2*>(cj + i), q1 = *reinterpret_cast<double2*>(cj + i + 4);
            double2 l0[FB], l1[FB];
#pragma unroll
            for (int p = 0; p < FB; ++p) {
                l0[p] = *reinterpret_cast<const double2*>(pk + p * FCS + i);
                l1[p] = *reinterpret_cast<const double2*>(pk + p * FCS + i + 4);
            }
#pragma unroll
            for (int p = 0; p < FB; ++p) {
                q0.x = fma(l0[p].x, m[p], q0.x); q0.y = fma(l0[p].y, m[p], q0.y);
                q1.x = fma(l1[p].x, m[p], q1.x); q1.y = fma(l1[p].y, m[p], q1.y);
            }
            *reinterpret_cast<double2*>(cj + i) = q0;
            *reinterpret_cast<double2*>(cj + i + 4) = q1;
        }
        if (i < SB) {
            double2 q0 = *reinterpret_cast<double2*>(cj + i);
#pragma unroll
            for (int p = 0; p < FB; ++p) {
                const double2 l0 = *reinterpret_cast<const double2*>(pk + p * FCS + i);
                q0.x = fma(l0.x, m[p], q0.x);
                q0.y = fma(l0.y, m[p], q0.y);
            }
            *reinterpret_cast<double2*>(cj + i) = q0;
        }
    }
    __syncthreads();
    BCR_MARK(3);
    if (tid < SB) sdi[tid] = sqrt(fmax(sinv[tid], 0.0));
    __syncthreads();
    // ---- write-out: factor in place of D_e, P, z (and Q in the one-pass kernel)
    for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
        const int r = i / SB, cc = i - r * SB;
        if (slice == 0) R[(size_t)e * SB2 + i] = r > cc ? sm[cc * FCS + r] * sinv[cc] : (r == cc ? sdi[r] : 0.0);
        if (a >= 0) (slice == 1 ? Q : P)[(size_t)e * SB2 + i] = sm[(SB + cc) * FCS + r] * sdi[r];
        if (!TWO_PASS && c >= 0) Q[(size_t)e * SB2 + i] = sm[(2 * SB + cc) * FCS + r] * sdi[r];
    }
    if (slice == 0 && tid < SB) rhs[(size_t)e * SB + tid] = sm[ZCOL * FCS + tid] * sdi[tid];
    BCR_MARK(4);
    if (!TWO_PASS || c < 0 || split) return;

    // ---- pass 2: Q = R^-1 Lc_c^T.  The right-hand buffer is reloaded with Lc_c^T; every column replays the
    //      elimination on its own (multipliers c_pq, 1/a_pp and the final panel columns are all in shared
    //      memory): two threads per column, __syncwarp only
    __syncthreads();                                   // the write-out above has read the buffer
    double* chunk = sm + SB * FCS;
    {
        const double* gC = Lc + (size_t)c * SB2;
        for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
            const int r = i / SB, cc = i - r * SB;
            cp_async8(&chunk[r * FCS + cc], gC + i);                           // Lc_c^T (cc, r) = Lc_c (r, cc)
        }
        if (tid < SB) chunk[tid * FCS + SB] = 0.0;
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    if (col < SB) {
        const unsigned mask = __activemask();          // whole warps except the last (150 threads)
        double* cq = chunk + col * FCS;
        for (int k0 = 0; k0 < SB; k0 += FB) {
            double u[FB], m[FB];
#pragma unroll
            for (int p = 0; p < FB; ++p) u[p] = cq[k0 + p];
#pragma unroll
            for (int p = 1; p < FB; ++p)
#pragma unroll
                for (int q = 0; q < p; ++q) u[p] = fma(-scpq_all[TWO_PASS ? k0 / FB : 0][p][q], u[q], u[p]);
            __syncwarp(mask);                          // both threads of the column have read the pivot rows
            if (half == 0) {
#pragma unroll
                for (int p = 1; p < FB; ++p) cq[k0 + p] = u[p];
            }
#pragma unroll
            for (int p = 0; p < FB; ++p) m[p] = -(u[p] * sinv[k0 + p]);
            const double* pk = sm + k0 * FCS;
            int i = k0 + FB;
            if (i & 1) {
                if (half == 0 && i < SB) {
                    double v = cq[i];
#pragma unroll
                    for (int p = 0; p < FB; ++p) v = fma(pk[p * FCS + i], m[p], v);
                    cq[i] = v;
                }
                ++i;
            }
            i += 2 * half;
            for (; i < SB; i += 4) {
                double2 q0 = *reinterpret_cast<double2*>(cq + i);
#pragma unroll
                for (int p = 0; p < FB; ++p) {
                    const double2 l0 = *reinterpret_cast<const double2*>(pk + p * FCS + i);
                    q0.x = fma(l0.x, m[p], q0.x);
                    q0.y = fma(l0.y, m[p], q0.y);
                }
                *reinterpret_cast<double2*>(cq + i) = q0;
            }
            __syncwarp(mask);                          // the next step's pivot rows may belong to the other thread
        }
    }
    __syncthreads();
    for (int i = tid; i < SB * SB; i += FACTOR_THREADS) {
        const int r = i / SB, cc = i - r * SB;
        Q[(size_t)e * SB2 + i] = chunk[cc * FCS + r] * sdi[r];
    }
}

#ifdef ACINO_BCR_TIMING
extern "C" void acino_debug_bcr_cycles(long long* out8) { cudaMemcpyFromSymbol(out8, g_bcr_cycles, sizeof(long long) * 8); }
extern "C" void acino_debug_bcr_reset() { long long z[8] = {0}; cudaMemcpyToSymbol(g_bcr_cycles, z, sizeof(z)); }
#endif

// ---- bcr_update: Schur update of one surviving super-block ------------------------------------------
//   Lc_j = -Q_el^T P_el                        (75 x 75 GEMM, warps 0-7: 15 x 15 grid of 5x5 register tiles)
//   D_j -= Q_el^T Q_el + P_er^T P_er           (two SYRKs: only the 120 upper tiles of each; warps 8-11 take
//                                               Q_el^T Q_el, warps 12-15 P_er^T P_er, summed through shared
//                                               memory, written with their mirror images)
//   b_j -= Q_el^T z_el + P_er^T z_er
// The three operand blocks are staged with 8-byte async copies issued together.
constexpr int UPDATE_THREADS = 512;
constexpr size_t BCR_UPDATE_SMEM = (size_t)(3 * SB * LD + 2 * SB) * sizeof(double);

struct TriTiles {
    unsigned char ty[120], tx[120];
};
constexpr TriTiles make_tri_tiles() {
    TriTiles t{};
    int n = 0;
    for (int y = 0; y < 15; ++y)
        for (int x = y; x < 15; ++x) {
            t.ty[n] = (unsigned char)y;
            t.tx[n] = (unsigned char)x;
            ++n;
        }
    return t;
}
__constant__ TriTiles c_tri_tiles = make_tri_tiles();

// acc(5x5) = A[:, i0:i0+5]^T B[:, j0:j0+5], A and B 75 x 75 row-major in shared memory (ld = LD)
__device__ __forceinline__ void tile_atb(const double* __restrict__ sA, const double* __restrict__ sB, const int i0,
                                         const int j0, double acc[5][5]) {
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[r][c] = 0.0;
#pragma unroll 3
    for (int k = 0; k < SB; ++k) {
        double a[5], b[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) a[r] = sA[k * LD + i0 + r];
#pragma unroll
        for (int c = 0; c < 5; ++c) b[c] = sB[k * LD + j0 + c];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
}

__global__ void __launch_bounds__(UPDATE_THREADS)
bcr_update_kernel(const int* __restrict__ surv /*[ns][3]*/, double* __restrict__ D, double* __restrict__ Lc,
                  const double* __restrict__ P, const double* __restrict__ Q, double* __restrict__ rhs) {
    extern __shared__ __align__(16) double sm[];
    double* sQ = sm;                       // Q_el
    double* sP = sm + SB * LD;             // P_el; after the GEMM: staging of the P_er^T P_er tiles
    double* sP2 = sm + 2 * SB * LD;        // P_er
    double* szl = sm + 3 * SB * LD;        // z_el
    double* szr = szl + SB;                // z_er
    const int j = surv[3 * blockIdx.x], el = surv[3 * blockIdx.x + 1], er = surv[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    for (int i = tid; i < SB * SB; i += UPDATE_THREADS) {
        const int r = i / SB, c = i - r * SB;
        if (el >= 0) {
            cp_async8(&sQ[r * LD + c], Q + (size_t)el * SB2 + i);
            cp_async8(&sP[r * LD + c], P + (size_t)el * SB2 + i);
        }
        if (er >= 0) cp_async8(&sP2[r * LD + c], P + (size_t)er * SB2 + i);
    }
    if (tid < SB) {
        szl[tid] = el >= 0 ? rhs[(size_t)el * SB + tid] : 0.0;
        szr[tid] = er >= 0 ? rhs[(size_t)er * SB + tid] : 0.0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    double acc[5][5];
    if (tid < 256) {
        // ---- Lc_j = -Q_el^T P_el
        const int ty = tid >> 4, tx = tid & 15;
        if (el >= 0 && ty < 15 && tx < 15) {
            tile_atb(sQ, sP, 5 * ty, 5 * tx, acc);
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) Lc[(size_t)j * SB2 + (5 * ty + r) * SB + 5 * tx + c] = -acc[r][c];
        }
        // ---- b_j -= Q_el^T z_el + P_er^T z_er
        if (tid < SB) {
            double s = 0.0;
            if (el >= 0)
                for (int k = 0; k < SB; ++k) s = fma(sQ[k * LD + tid], szl[k], s);
            if (er >= 0)
                for (int k = 0; k < SB; ++k) s = fma(sP2[k * LD + tid], szr[k], s);
            rhs[(size_t)j * SB + tid] -= s;
        }
        asm volatile("bar.sync 2, 512;" ::: "memory");      // P_el is free from here on
        asm volatile("bar.sync 3, 512;" ::: "memory");      // staged tiles are complete
    } else {
        // ---- D_j -= Q_el^T Q_el + P_er^T P_er, upper tiles; group A (Q) = threads 256..383, group B (P) = 384..511
        const int t = (tid - 256) & 127;
        const bool grpB = tid >= 384;
        const bool have = t < 120 && (grpB ? er >= 0 : el >= 0);
        int ty = 0, tx = 0;
        if (t < 120) {
            ty = c_tri_tiles.ty[t];
            tx = c_tri_tiles.tx[t];
        }
        if (have) {
            const double* M = grpB ? sP2 : sQ;
            tile_atb(M, M, 5 * ty, 5 * tx, acc);
        } else {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) acc[r][c] = 0.0;
        }
        asm volatile("bar.sync 2, 512;" ::: "memory");      // the GEMM no longer reads P_el
        if (grpB && t < 120) {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) sP[(5 * ty + r) * LD + 5 * tx + c] = acc[r][c];
        }
        asm volatile("bar.sync 3, 512;" ::: "memory");
        if (!grpB && t < 120) {
            // read-modify-write of this thread's own 25 entries: all loads first (a store to D followed by a load
            // from D would otherwise be serialised - one global round trip per entry); the mirror image gets the
            // same value (D stays exactly symmetric)
            double* Dj = D + (size_t)j * SB2;
            double nv[5][5];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) nv[r][c] = Dj[(5 * ty + r) * SB + 5 * tx + c];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) nv[r][c] -= acc[r][c] + sP[(5 * ty + r) * LD + 5 * tx + c];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    Dj[(5 * ty + r) * SB + 5 * tx + c] = nv[r][c];
                    if (ty != tx) Dj[(5 * tx + c) * SB + 5 * ty + r] = nv[r][c];
                }
        }
    }
}

// Two-stage variant for the large levels (more surviving blocks than SMs): two operand buffers instead of three
// (92 KB) and 256 threads with up to 128 registers -> two CTAs per SM.
//   stage 1: buf0 = Q_el, buf1 = P_er: the two SYRKs in parallel (threads 0..119 / 128..247), D_j updated by the
//            P_er group first, then (after a barrier) by the Q_el group; b_j
//   stage 2: buf1 <- P_el: the GEMM Lc_j = -Q_el^T P_el on 225 threads
constexpr int UPDATE2_THREADS = 256;
constexpr size_t BCR_UPDATE2_SMEM = (size_t)(2 * SB * LD + 2 * SB) * sizeof(double);

__device__ __forceinline__ void rmw_tile(double* __restrict__ Dj, const int ty, const int tx, const double acc[5][5]) {
    double nv[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) nv[r][c] = Dj[(5 * ty + r) * SB + 5 * tx + c];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            nv[r][c] -= acc[r][c];
            Dj[(5 * ty + r) * SB + 5 * tx + c] = nv[r][c];
            if (ty != tx) Dj[(5 * tx + c) * SB + 5 * ty + r] = nv[r][c];
        }
}

__global__ void __launch_bounds__(UPDATE2_THREADS, 2)
bcr_update2_kernel(const int* __restrict__ surv /*[ns][3]*/, double* __restrict__ D, double* __restrict__ Lc,
                   const double* __restrict__ P, const double* __restrict__ Q, double* __restrict__ rhs, const int split) {
    extern __shared__ __align__(16) double sm[];
    double* b0 = sm;                       // Q_el
    double* b1 = sm + SB * LD;             // P_er, then P_el
    double* szl = sm + 2 * SB * LD;        // z_el
    double* szr = szl + SB;                // z_er
    const int j = surv[3 * blockIdx.x], el = surv[3 * blockIdx.x + 1], er = surv[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    // Split mode (small levels): CTA (., 0) does stage 1 (the SYRKs onto D_j and b_j), CTA (., 1) stage 2 (the GEMM that
    // gives Lc_j) - disjoint outputs, so the two halves of a block's update run on two SMs at once.
    const int part = split ? (int)blockIdx.y : -1;
    if (part == 1) {
        if (el < 0) return;
        for (int i = tid; i < SB * SB; i += UPDATE2_THREADS) {
            const int r = i / SB, c = i - r * SB;
            cp_async8(&b0[r * LD + c], Q + (size_t)el * SB2 + i);
            cp_async8(&b1[r * LD + c], P + (size_t)el * SB2 + i);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        const int gy = tid >> 4, gx = tid & 15;
        if (gy < 15 && gx < 15) {
            double acc[5][5];
            tile_atb(b0, b1, 5 * gy, 5 * gx, acc);
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) Lc[(size_t)j * SB2 + (5 * gy + r) * SB + 5 * gx + c] = -acc[r][c];
        }
        return;
    }
    for (int i = tid; i < SB * SB; i += UPDATE2_THREADS) {
        const int r = i / SB, c = i - r * SB;
        if (el >= 0) cp_async8(&b0[r * LD + c], Q + (size_t)el * SB2 + i);
        if (er >= 0) cp_async8(&b1[r * LD + c], P + (size_t)er * SB2 + i);
    }
    if (tid < SB) {
        szl[tid] = el >= 0 ? rhs[(size_t)el * SB + tid] : 0.0;
        szr[tid] = er >= 0 ? rhs[(size_t)er * SB + tid] : 0.0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    double acc[5][5];
    double* Dj = D + (size_t)j * SB2;
    // ---- stage 1: SYRKs (upper tiles), group A = threads 0..127 (Q_el), group B = 128..255 (P_er)
    const int t = tid & 127;
    const bool grpB = tid >= 128;
    const bool have = t < 120 && (grpB ? er >= 0 : el >= 0);
    int ty = 0, tx = 0;
    if (t < 120) {
        ty = c_tri_tiles.ty[t];
        tx = c_tri_tiles.tx[t];
    }
    if (have) {
        const double* M = grpB ? b1 : b0;
        tile_atb(M, M, 5 * ty, 5 * tx, acc);
        if (grpB) rmw_tile(Dj, ty, tx, acc);
    }
    // b_j -= Q_el^T z_el + P_er^T z_er   (threads 120..127 and 248..255 are idle above: any thread may do it after its tile)
    if (tid < SB) {
        double sacc = 0.0;
        if (el >= 0)
            for (int k = 0; k < SB; ++k) sacc = fma(b0[k * LD + tid], szl[k], sacc);
        if (er >= 0)
            for (int k = 0; k < SB; ++k) sacc = fma(b1[k * LD + tid], szr[k], sacc);
        rhs[(size_t)j * SB + tid] -= sacc;
    }
    __syncthreads();                       // group B's update of D_j is visible; P_er no longer needed
    if (have && !grpB) rmw_tile(Dj, ty, tx, acc);
    if (el < 0 || part == 0) return;
    // ---- stage 2: Lc_j = -Q_el^T P_el
    for (int i = tid; i < SB * SB; i += UPDATE2_THREADS) {
        const int r = i / SB, c = i - r * SB;
        cp_async8(&b1[r * LD + c], P + (size_t)el * SB2 + i);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int gy = tid >> 4, gx = tid & 15;
    if (gy < 15 && gx < 15) {
        tile_atb(b0, b1, 5 * gy, 5 * gx, acc);
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) Lc[(size_t)j * SB2 + (5 * gy + r) * SB + 5 * gx + c] = -acc[r][c];
    }
}

// x_e = R^-T (z - P x_a - Q x_c), R = L Delta^1/2 stored by bcr_factor in D_e.  256 threads.
//   1. v = z - [P | Q] [x_a ; x_c]: warp per row, lanes across the row (coalesced), 30 independent loads in flight per
//      thread; L is staged in shared memory meanwhile
//   2. backward substitution L^T x = Delta^-1/2 v by ONE warp, no barriers: lane holds rows lane, lane+32,
//      lane+64; once x_k is final it is broadcast by shuffle and u_i -= L_ki x_k for i < k (row k of L).
__global__ void __launch_bounds__(256)
bcr_backsub_kernel(const int* __restrict__ elim, const double* __restrict__ D /*factor*/, const double* __restrict__ P,
                   const double* __restrict__ Q, const double* __restrict__ rhs, double* __restrict__ x) {
    __shared__ __align__(16) double sxx[2 * SB + 2];       // [x_a ; x_c]
    __shared__ double spart[1][SB];
    __shared__ double sL[SB * LD];
    const int e = elim[3 * blockIdx.x], a = elim[3 * blockIdx.x + 1], c = elim[3 * blockIdx.x + 2];
    const int tid = threadIdx.x;
    if (tid < SB) sxx[tid] = a >= 0 ? x[(size_t)a * SB + tid] : 0.0;
    else if (tid < 2 * SB) sxx[tid] = c >= 0 ? x[(size_t)c * SB + tid - SB] : 0.0;
    for (int i = tid; i < SB * SB; i += 256) {      // L: async, lands while the mat-vec below runs
        const int r = i / SB;
        cp_async8(&sL[r * LD + (i - r * SB)], D + (size_t)e * SB2 + i);
    }
    __syncthreads();
    {
        // v_r = sum_c P[r][c] x_a[c] + Q[r][c] x_c[c]: warp per row, lanes across the row (coalesced 600-byte rows, all
        // loads of a warp's 9-10 rows independent), fixed-order butterfly sum.  spart[0] holds the result.
        const int warp = tid >> 5, lane = tid & 31;
        const double* Pe = P + (size_t)e * SB2;
        const double* Qe = Q + (size_t)e * SB2;
        // rows r = warp + 8 i, i = 0..9 in two batches of five: 30 independent loads in flight per thread
#pragma unroll
        for (int batch = 0; batch < 2; ++batch) {
            double pv[5][3], qv[5][3];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int r = warp + 8 * (5 * batch + i);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int cc = lane + 32 * k;
                    const bool ok = r < SB && cc < SB;
                    pv[i][k] = (ok && a >= 0) ? Pe[r * SB + cc] : 0.0;
                    qv[i][k] = (ok && c >= 0) ? Qe[r * SB + cc] : 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int r = warp + 8 * (5 * batch + i);
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int cc = lane + 32 * k;
                    if (cc < SB) s = fma(qv[i][k], sxx[SB + cc], fma(pv[i][k], sxx[cc], s));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && r < SB) spart[0][r] = s;
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        double u[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = lane + 32 * j;
            u[j] = i < SB ? (rhs[(size_t)e * SB + i] - spart[0][i]) * sL[i * LD + i] : 0.0;
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


// Levels with at most SPLIT_MAX blocks run two CTAs per block (split kernels above): the GPU has 148 SMs and these levels
// are bound by the latency of one block.  Levels with more blocks than SMs use the two-pass / two-stage kernels (two CTAs
// per SM); in between, the one-pass kernels.
constexpr int SPLIT_MAX = 74;
constexpr int TWO_PASS_FROM = 149;

static cudaError_t bcr_configure() {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || configured[dev]) return cudaSuccess;        // the attribute is per device
    cudaError_t e = cudaFuncSetAttribute(bcr_factor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_FACTOR_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bcr_factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_FACTOR2_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bcr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_UPDATE_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bcr_update2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCR_UPDATE2_SMEM);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
    return cudaSuccess;
}

cudaError_t launch_bcr_factor(int n_elim, const int* elim, const double* D, const double* Lc, double* P, double* Q,
                              double* R, double* rhs, int* info, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    cudaError_t e = bcr_configure();
    if (e != cudaSuccess) return e;
    if (n_elim <= SPLIT_MAX)
        bcr_factor_kernel<true><<<dim3(n_elim, 2), FACTOR_THREADS, BCR_FACTOR2_SMEM, s>>>(elim, D, Lc, P, Q, R, rhs, info, 1);
    else if (n_elim >= TWO_PASS_FROM)
        bcr_factor_kernel<true><<<n_elim, FACTOR_THREADS, BCR_FACTOR2_SMEM, s>>>(elim, D, Lc, P, Q, R, rhs, info, 0);
    else
        bcr_factor_kernel<false><<<n_elim, FACTOR_THREADS, BCR_FACTOR_SMEM, s>>>(elim, D, Lc, P, Q, R, rhs, info, 0);
    return cudaGetLastError();
}

cudaError_t launch_bcr_update(int n_surv, const int* surv, double* D, double* Lc, const double* P, const double* Q,
                              double* rhs, cudaStream_t s) {
    if (n_surv <= 0) return cudaSuccess;
    cudaError_t e = bcr_configure();
    if (e != cudaSuccess) return e;
    if (n_surv <= SPLIT_MAX)
        bcr_update2_kernel<<<dim3(n_surv, 2), UPDATE2_THREADS, BCR_UPDATE2_SMEM, s>>>(surv, D, Lc, P, Q, rhs, 1);
    else if (n_surv >= TWO_PASS_FROM)
        bcr_update2_kernel<<<n_surv, UPDATE2_THREADS, BCR_UPDATE2_SMEM, s>>>(surv, D, Lc, P, Q, rhs, 0);
    else
        bcr_update_kernel<<<n_surv, UPDATE_THREADS, BCR_UPDATE_SMEM, s>>>(surv, D, Lc, P, Q, rhs);
    return cudaGetLastError();
}

cudaError_t launch_bcr_backsub(int n_elim, const int* elim, const double* R, const double* P, const double* Q,
                               const double* rhs, double* x, cudaStream_t s) {
    if (n_elim <= 0) return cudaSuccess;
    bcr_backsub_kernel<<<n_elim, 256, 0, s>>>(elim, R, P, Q, rhs, x);
    return cudaGetLastError();
}

}  // namespace acino
