// fte_eval: per-frame reprojection cost, gradient and Gauss-Newton block of the FTE objective.
//
// Replaces (reference, /root/reference/src/all_optimizations.py): the SymPy FK chain :66-179,
// pose_constraint :359-365, measurement_constraints :394-399 through pt3d_to_2d :193-209, the
// measurement term of obj :494-497 with redescending_loss (build.py:382-395), and the
// automatic differentiation IPOPT's ASL performs on those expression trees each iteration.
//
// Layout / mapping (B200, CUDA cores only - the contraction is tiny, irregular and sparse):
//   persistent CTAs (one wave, 4 per SM), one tile = FT frames = FT * 20 threads; tiles are drawn from a global ticket
//   counter (the CTAs of an SM do not run at the same speed).  Phases, separated by __syncthreads():
//   P1a sin / cos     thread <-> (angle, frame)
//   P1b FK            3 threads <-> frame        one thread per ROW of the rotation chain (right-multiplications keep the
//                                                rows independent), on three warps (trunk + tail / front legs / back legs):
//                                                component i of every marker position and rotation axis
//   P2  projection    thread <-> (frame, marker) loops over camera pairs (unrolled for six cameras): fisheye projection, 2x3
//                                                Jacobian, redescending loss; accumulates the
//                                                marker's 3x3 normal block A_l, 3-vector b_l and
//                                                cost in registers (no cross-thread reduction),
//                                                then forms the marker's 6x6 "spatial inertia"
//                                                B_l^T A_l B_l and wrench B_l^T b_l, B_l=[-[p]x I]
//   P3  subtree sums  thread <-> (frame, comp)   composite-rigid-body style accumulation of the
//                                                27 components up the kinematic tree (fixed order)
//   P4  columns       thread <-> (frame, angle)  y = I_subtree tau_beta in registers; H[a][b] = tau_a . y for
//                                                every ancestor angle a; g[b] = tau_b . wrench
//   P5  write-out     bulk async stores (TMA) of {cost, g[25], H[325]} from shared memory; the next tile's inputs are
//                     requested before / during it and land while this tile is still being reduced
// d p_l / d angle = omega x (p_l - pivot) = B_l tau (twist about the head point), so
// H = sum_l J_l^T A_l J_l collapses to tau_a^T I_{deeper subtree} tau_b: ~4 kFMA instead of
// ~8 kFMA per frame for the chain rule and no 60x25 Jacobian is ever materialised.
#include <cstdlib>
#include <type_traits>

#include "acino_common.cuh"
#include "cheetah_fk.cuh"

namespace acino {

constexpr int N_PAIR = NANG * (NANG + 1) / 2;  // 253 unordered angle pairs (incl. the diagonal)

// Column table of the H assembly (P4).  One thread owns one angle slot `be` of one frame: it forms
// y = I_subtree(joint(be)) tau_be in registers and from it every entry H[al][be] = tau_al . y with `al` an
// ancestor-or-self angle of `be` (same joint: al <= be) - column `be` of the block.  Ancestors at the head and neck joints
// ("trunk", 6 slots, 117 of the 185 related pairs) pivot on the head point: v_al = 0, so their entries are 3-term dot
// products omega_al . y_top with omega_al loaded once per frame (broadcast).  The other ancestors (<= 8) come from a
// per-slot list.  Unrelated pairs (disjoint subtrees, 68) are structural zeros: a list of their H indices, written without
// arithmetic.  Task order: the 16 columns that have non-trunk ancestors, sorted by list length so that the four columns
// sharing a warp do similar work; then the six trunk columns (empty lists, v = 0: a shorter code path of their own).
constexpr int N_TRUNK = 6;
constexpr int k_trunk_slot[N_TRUNK] = {0, 3, 17, 1, 4, 18};     // phi0 theta0 psi0 | phi1 theta1 psi1
constexpr int trunk_slot(int k) { return k == 0 ? 0 : k == 1 ? 3 : k == 2 ? 17 : k == 3 ? 1 : k == 4 ? 4 : 18; }
constexpr int MAX_ANC = 8;
constexpr int N_GENERIC = NANG - N_TRUNK;      // 16 columns with non-trunk ancestors-or-self
#ifndef ACINO_UNROLL_PAIRS
#define ACINO_UNROLL_PAIRS 3             // the camera loop is unrolled for this many camera pairs (the reference's six cameras)
#endif
constexpr int N_ZERO = N_PAIR - 185;
constexpr unsigned NO_ENTRY = 0xFFFFu;
struct ColEntry {                       // 112 bytes of ready-to-use 32-bit words: seven 16-byte loads, no field extraction
    unsigned tau_off, ij_off, n_anc, slot;   // be * TAU_STRIDE, joint * (NSP + 1) (float offsets), list length, be
    unsigned trunk_idx[N_TRUNK];        // packed-H index of (trunk slot k, be); NO_ENTRY: not a related pair in this order
    unsigned pad0[2];
    unsigned anc_off[MAX_ANC];          // non-trunk ancestor-or-self slots as float offsets al * TAU_STRIDE (unused: 0)
    unsigned anc_idx[MAX_ANC];          // their packed-H indices
};
static_assert(sizeof(ColEntry) == 112, "ColEntry layout");
struct ColTable {
    ColEntry col[NANG];
    unsigned short zero_idx[N_ZERO + 4];
    int n_rel;
};
constexpr bool is_trunk_slot(int a) {
    for (int k = 0; k < N_TRUNK; ++k)
        if (k_trunk_slot[k] == a) return true;
    return false;
}
constexpr bool related(int al, int be) {     // H[al][be] is computed as column `be`
    const int ja = k_angle_joint[al], jb = k_angle_joint[be];
    return joint_is_anc(ja, jb) && (ja != jb || al <= be);
}
constexpr unsigned short packed_index(int al, int be) {
    const int sa = 3 + al, sb = 3 + be;
    const int lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
    return (unsigned short)(lo * NA - (lo * (lo - 1)) / 2 + (hi - lo));
}
constexpr ColTable make_col_table() {
    ColTable t{};
    int n_anc[NANG] = {};
    for (int be = 0; be < NANG; ++be)
        for (int al = 0; al < NANG; ++al)
            if (related(al, be) && !is_trunk_slot(al)) ++n_anc[be];
    // slots sorted by list length (stable)
    int order[NANG] = {};
    for (int i = 0; i < NANG; ++i) order[i] = i;
    for (int i = 1; i < NANG; ++i)
        for (int j = i; j > 0 && n_anc[order[j - 1]] > n_anc[order[j]]; --j) {
            const int tmp = order[j];
            order[j] = order[j - 1];
            order[j - 1] = tmp;
        }
    int n_rel = 0;
    for (int q = 0; q < NANG; ++q) {
        // tasks 0 .. 15: the columns with non-trunk ancestors, lightest first (the six trunk slots have empty lists and sort
        // first); tasks 16 .. 21: the trunk columns
        const int be = order[q < N_GENERIC ? q + N_TRUNK : q - N_GENERIC];
        ColEntry& c = t.col[q];
        c.slot = (unsigned)be;
        c.tau_off = (unsigned)(be * TAU_STRIDE);
        c.ij_off = (unsigned)(k_angle_joint[be] * (NSP + 1));
        for (int k = 0; k < N_TRUNK; ++k) {
            const bool on = related(k_trunk_slot[k], be);
            c.trunk_idx[k] = on ? packed_index(k_trunk_slot[k], be) : NO_ENTRY;
            n_rel += on;
        }
        int n = 0;
        for (int al = 0; al < NANG; ++al)
            if (related(al, be) && !is_trunk_slot(al)) {
                c.anc_off[n] = (unsigned)(al * TAU_STRIDE);
                c.anc_idx[n] = packed_index(al, be);
                ++n;
            }
        c.n_anc = (unsigned)n;
        n_rel += n;
    }
    int nz = 0;
    for (int be = 0; be < NANG; ++be)
        for (int al = 0; al < be; ++al)
            if (!related(al, be) && !related(be, al)) t.zero_idx[nz++] = packed_index(al, be);
    t.n_rel = n_rel * 1000 + nz;
    return t;
}
constexpr int max_anc_len() {
    int m = 0;
    for (int be = 0; be < NANG; ++be) {
        int n = 0;
        for (int al = 0; al < NANG; ++al) n += related(al, be) && !is_trunk_slot(al);
        m = n > m ? n : m;
    }
    return m;
}
static_assert(max_anc_len() == MAX_ANC, "kinematic tree changed: check the column table");
constexpr bool trunk_tasks_ok() {
    const ColTable t = make_col_table();
    for (int q = 0; q < NANG; ++q) {
        const bool trunk = is_trunk_slot((int)t.col[q].slot);
        if (trunk != (q >= N_GENERIC) || (trunk && (t.col[q].n_anc != 0 || k_angle_pivot[t.col[q].slot] >= 0))) return false;
        if (!trunk)
            for (int k = 0; k < N_TRUNK; ++k)
                if (t.col[q].trunk_idx[k] == NO_ENTRY) return false;
    }
    return true;
}
static_assert(trunk_tasks_ok(), "tasks 16..21 must be the trunk columns (head-point pivot, no non-trunk ancestor), and every trunk "
                                "slot an ancestor of tasks 0..15");
static_assert(make_col_table().n_rel == 185 * 1000 + N_ZERO, "kinematic tree changed: 185 related + 68 unrelated pairs");
__constant__ ColTable c_col = make_col_table();
__device__ __align__(16) const ColTable d_col = make_col_table();     // global-memory copy: source of the bulk (TMA) copy
constexpr unsigned COL_BYTES = sizeof(ColEntry) * NANG + sizeof(unsigned short) * (N_ZERO + 4);   // 2464 + 144
static_assert(COL_BYTES % 16 == 0, "bulk copy size");

// the 16 angle slots whose rotation does not pivot on the head point: (slot, pivot marker) - v = pivot x omega is formed
// from the stored components by otherwise idle threads of the subtree-sum phase
struct PivotTable {
    unsigned char slot[16], marker[16];
};
constexpr PivotTable make_pivot_table() {
    PivotTable t{};
    int n = 0;
    for (int a = 0; a < NANG; ++a)
        if (k_angle_pivot[a] >= 0) {
            t.slot[n] = (unsigned char)a;
            t.marker[n] = (unsigned char)k_angle_pivot[a];
            ++n;
        }
    return t;
}
constexpr int count_pivoted() {
    int n = 0;
    for (int a = 0; a < NANG; ++a) n += k_angle_pivot[a] >= 0;
    return n;
}
static_assert(count_pivoted() == 16, "kinematic tree changed: check the pivot table");
constexpr unsigned pivoted_mask() {
    unsigned m = 0;
    for (int a = 0; a < NANG; ++a) m |= (k_angle_pivot[a] >= 0 ? 1u : 0u) << a;
    return m;
}
constexpr unsigned k_pivoted_mask = pivoted_mask();     // bit a: angle slot a pivots on a marker (v != 0)
__constant__ PivotTable c_piv = make_pivot_table();

// bulk-copied input tiles: the measurement and weight rows of the FT frames of a tile are contiguous in global memory,
// so each tile is ONE bulk copy (issuing a copy costs the issuing thread ~100 cycles: 2 copies per tile, not 2 per frame)
template <int FT, int MAXC>
struct InTiles {
    float2 meas[FT * MAXC * NL];          // [FT][C][NL] (u,v)
    float w[FT * MAXC * NL];              // [FT][C][NL]
};
#ifndef ACINO_MAXC_STAGE
#define ACINO_MAXC_STAGE (ACINO_MAX_CAMS / 2)
#endif
constexpr int MAXC_STAGE = ACINO_MAXC_STAGE;     // cameras the staged input tiles hold; more cameras: direct loads

// Shared memory of one CTA (persistent: the CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...).  Two regions are
// re-used inside a tile, arranged so that the NEXT tile's inputs can be fetched while this tile is still being reduced:
//   region A: per-marker inertia Il (P2 -> P3), then the staged outputs o (P4 -> bulk stores of P5, read by the TMA
//             engine while the next tile's FK runs)
//   region B: this tile's measurement / weight tiles (TMA -> P2), then the per-joint subtree sums Ij (P3 -> P4a); free
//             again after P4a, when the next tile's bulk copies are issued (they land during P4b, P5 and the next FK)
template <int FT>
struct __align__(16) Smem {
    float x[2][FT][NA];                // state, double buffered (next tile prefetched)
    float p[FT][NL][3];                // marker positions relative to the head point
    float2 sc[FT][NANG];               // (sin, cos) of every angle
    __align__(16) float costp[FT][NL]; // per-(frame, marker) cost partials
    __align__(16) ColEntry col[NANG];  // column table of the H assembly (copy of c_col), followed by the zero list
    unsigned short zero_idx[N_ZERO + 4];
    __align__(16) float tau[FT][TAUF]; // (omega, v = pivot x omega) per angle, stride 8
    unsigned long long mbar[2];        // [0] state tile landed, [1] measurement + weight tiles landed
    int next_tile;                     // dynamic schedule: the tile after this CTA's next one (published by the claiming thread)
    union {                            // region A
        // per-marker spatial inertia + wrench, 27 padded to 28: STS.128 (P2) / LDS.64 (P3).  Frame stride 572 words =
        // 28 (mod 32): in P3 a warp covers 14 + 14 + 4 (component pair, frame) lanes and the lanes of the next frame
        // continue in the banks where the previous frame's stop - with the natural stride 560 = 16 (mod 32) every one of
        // P3's loads was a 2-way conflict
        __align__(16) float Il[FT][NL * (NSP + 1) + 12];
        struct {                       // staged outputs in their global layout
            float H[FT][NU];
            float g[FT][NA];
            float cost[FT];
        } o;
    };
    union {                            // region B
        InTiles<FT, MAXC_STAGE> in;
        // subtree spatial inertia + wrench per joint (27 padded to 28: float4 loads); frame stride 412 words = 28 (mod 32):
        // the 8 frames of a quarter-warp hit 8 distinct 4-bank groups in P4's LDS.128 (392 gave 2-way conflicts there) and
        // P3's STS.64 of neighbouring frames do not collide either (396 = 12 (mod 32) did)
        __align__(16) float Ij[FT][NJ * (NSP + 1) + 20];
    };
};

// ---- bulk async copy (TMA, 1-D) + mbarrier helpers ------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk store (TMA, 1-D) of a staged output tile; the issuing thread waits until the
// source has been read before the CTA may retire
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

#ifdef ACINO_PHASE_TIMING
__device__ long long g_phase_cycles[16];
__device__ int g_phase_count;
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd((unsigned long long*)&g_phase_cycles[i], (unsigned long long)(_t - _tprev)); _tprev = _t; } } while (0)
// camera-loop intervals of the first 128 tiles of every CTA (SM clock: comparable between the CTAs of one SM) - scripts/p2_overlap.py
__device__ long long g_p2_trace[1024 * 128 * 2];
__device__ int g_p2_smid[1024];
#define P2_TRACE(it, which) do { if (threadIdx.x == 0 && (it) < 128 && blockIdx.x < 1024) g_p2_trace[(blockIdx.x * 128 + (it)) * 2 + (which)] = clock64(); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#define P2_TRACE(it, which) do { } while (0)
#endif

// range-reduced MUFU sine / cosine: |error| < 4e-7 for any |x| up to a few hundred revolutions (the reduction is one
// fp32 FMA; the reference's angles are bounded by pi).  The same values feed cost, gradient and H.
__device__ __forceinline__ void fast_sincos(const float x, float& sn, float& cs) {
    const float xr = fmaf(-6.283185307179586f, rintf(x * 0.15915494309189535f), x);
    sn = __sinf(xr);
    cs = __cosf(xr);
}

template <int FT, bool WANT_H>
__device__ __forceinline__ void
fte_eval_body(const SceneF& scene, const int n_frames, const int use_bulk,
              const float* __restrict__ xg, const float* __restrict__ meas,
              const float* __restrict__ wts, float* __restrict__ cost_out,
              float* __restrict__ g_out, float* __restrict__ H_out, int* __restrict__ sched) {
    constexpr int NT = FT * NL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<FT>& S = *reinterpret_cast<Smem<FT>*>(smem_raw);
    const int tid = threadIdx.x;
    const int C = scene.n_cams;
    const int n_tiles = (n_frames + FT - 1) / FT;
    // the input tiles of a full tile are contiguous, 16-byte aligned blocks of global memory: stage them with 1-D
    // bulk async copies (TMA) issued one tile ahead
    const bool bulk_in = (use_bulk & 1) && C <= MAXC_STAGE;
    float2* const in_meas = S.in.meas;
    float* const in_w = S.in.w;
    const int ms2 = C * NL, ws = C * NL;            // frame strides of the staged tiles: float2 units / float units
    const unsigned bx = FT * NA * 4, bm = FT * C * NL * 8, bw = FT * C * NL * 4;     // bytes: state / meas / weight tile
    auto issue_x = [&](const int t, const int buf, const bool with_tab) {
        mbar_expect_tx(&S.mbar[0], bx + (with_tab ? COL_BYTES : 0));
        bulk_g2s(&S.x[buf][0][0], xg + (size_t)t * FT * NA, bx, &S.mbar[0]);
        if (with_tab) bulk_g2s(&S.col[0], &d_col.col[0], COL_BYTES, &S.mbar[0]);
    };
    auto issue_mw = [&](const int t) {
        mbar_expect_tx(&S.mbar[1], bm + bw);
        bulk_g2s(in_meas, meas + (size_t)t * FT * C * NL * 2, bm, &S.mbar[1]);
        bulk_g2s(in_w, wts + (size_t)t * FT * C * NL, bw, &S.mbar[1]);
    };
#ifdef ACINO_PHASE_TIMING
    long long _tprev = clock64();
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_p2_smid[blockIdx.x] = (int)smid;
    }
#endif

    // sin / cos of the 22 angles of tile t, one thread per (angle, frame); its state tile is in (or goes to) S.x[buf]
    unsigned ph_x = 0, ph_m = 0;       // completed phases of mbar[0] / mbar[1]
    auto sincos_tile = [&](const int t, const int buf) {
        const int tf0 = t * FT;
        const int tnf = min(FT, n_frames - tf0);
        float (*X)[NA] = S.x[buf];
        if (bulk_in && tnf == FT) {
            mbar_wait(&S.mbar[0], ph_x & 1);
            ++ph_x;
            for (int k = tid; k < FT * NANG; k += NT) {
                const int a = k / FT, f = k - a * FT;
                float sn, cs;
                fast_sincos(X[f][3 + a], sn, cs);
                S.sc[f][a] = make_float2(sn, cs);
            }
        } else {
            // partial tile or unaligned inputs: the state goes to shared memory by plain loads (read again by the
            // camera loop, several barriers from here); the angles are read straight from global memory
            for (int i = tid; i < FT * NA; i += NT) {
                const int f = i / NA;
                (&X[0][0])[i] = (f < tnf) ? xg[(size_t)tf0 * NA + i] : 0.f;
            }
            for (int k = tid; k < FT * NANG; k += NT) {
                const int a = k / FT, f = k - a * FT;
                float sn, cs;
                fast_sincos(f < tnf ? xg[(size_t)(tf0 + f) * NA + 3 + a] : 0.f, sn, cs);
                S.sc[f][a] = make_float2(sn, cs);
            }
        }
    };

    // ---- prologue: barriers, the constant zero entries of the twists, the first tile's copies, table and sin / cos
    int tile = blockIdx.x;
    const bool first_staged = bulk_in && (tile + 1) * FT <= n_frames;
    if (tid == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (first_staged) {
            issue_x(tile, 0, true);
            issue_mw(tile);
        }
    }
    // v = 0 for the angles that pivot on the head point, and the all-zero slot NANG (never written again)
    for (int i = tid; i < FT * (NANG + 1); i += NT) {
        const int f = i / (NANG + 1), a = i - f * (NANG + 1);
        if (a == NANG || !((k_pivoted_mask >> a) & 1u)) {
            float* t = &S.tau[f][a * TAU_STRIDE];
            t[3] = 0.f; t[4] = 0.f; t[5] = 0.f;
            if (a == NANG) { t[0] = 0.f; t[1] = 0.f; t[2] = 0.f; }
        }
    }
    if (!first_staged)          // (a staged first tile brings the table with its state)
        for (int i = tid; i < (int)(COL_BYTES / 4); i += NT)
            reinterpret_cast<unsigned*>(&S.col[0])[i] = reinterpret_cast<const unsigned*>(&c_col.col[0])[i];
    // The bulk copies are issued by threads of warps 3 and 4, which have nothing to do while warps 0-2 run the FK chain of the
    // next tile (issuing one costs its thread ~100 cycles: on warps 0 / 1 the FK chain started ~600 cycles late)
    constexpr int ST_TID = NT > 128 ? 128 : 0;     // output stores (and the wait for them)
    constexpr int MW_TID = NT > 97 ? 97 : 32;      // next tile's measurement / weight tiles (thread 96: next tile's state)
    bool out_pending = false;          // (thread ST_TID) bulk stores of the previous tile may still be reading region A
    __syncthreads();
    sincos_tile(tile, 0);              // P0 + P1a of the first tile
    __syncthreads();
    PHASE_MARK(1);

    // Tile schedule.  Static (tile, tile + grid, ...) when every CTA has one tile, or no counter was given.  Otherwise
    // dynamic: the first two tiles are blockIdx.x and blockIdx.x + grid, every further one is claimed from a global ticket
    // counter TWO tiles ahead: the atomic is issued during the FK phase and its result is first touched after the camera
    // loop (claimed one tile ahead, the ~1500-cycle round trip under load sat on the FK phase's critical path).  The CTAs
    // of a persistent wave do not run at the same speed - with the static schedule the fast ones had exited while the slow
    // ones still had tiles left (17.2 of 20 warps active on average); the last CTA to leave resets the counter pair for
    // the next launch.
    static_assert(NT > 128, "threads 96 / 97 / 128 issue the copies and draw the tickets");
#ifdef ACINO_STATIC_SCHED
    const bool dynamic = false;
#else
    const bool dynamic = sched != nullptr && n_tiles > (int)gridDim.x;
#endif
    int tile_next = tile + gridDim.x;
    int claimed = 0;                   // (thread 96) ticket drawn during this tile's FK phase
    for (int it = 0; tile < n_tiles; ++it) {
    const int f0 = tile * FT;
    const int nf = min(FT, n_frames - f0);
    const bool staged = bulk_in && nf == FT;
    const int xb = it & 1;
    float (*Sx)[NA] = S.x[xb];
    // (P0 + P1a, the sin / cos of this tile's angles, ran before the barrier that follows the previous tile's camera loop)

    // ---- P1b: rotation chain, three threads per frame (one per ROW of the chain: right-multiplications keep rows
    //      independent): component i of every marker position and rotation axis
    //      Warp 0: head / neck / torso / tail, warp 1: front legs, warp 2: back legs (each re-derives the part of the trunk
    //      it hangs from: cheaper than waiting for it).  Warp 3: the next tile's state (every thread passed its wait on
    //      mbar[0] during the previous tile, whose camera loop was the last reader of the other buffer)
    {
        const int wq = tid >> 5, ln = tid & 31;
        if (wq < 3) {
            if (ln < 3 * FT) {
                const int f = ln / 3, i = ln - 3 * f;
                if (wq == 0) cheetah_fk_row<1>(S.sc[f], i, &S.p[f][0][0], &S.tau[f][0]);
                else if (wq == 1) cheetah_fk_row<2>(S.sc[f], i, &S.p[f][0][0], &S.tau[f][0]);
                else cheetah_fk_row<4>(S.sc[f], i, &S.p[f][0][0], &S.tau[f][0]);
            }
        } else if (tid == 96) {
            if (dynamic) claimed = atomicAdd(sched, 1);
            if (bulk_in && tile_next < n_tiles && (tile_next + 1) * FT <= n_frames) issue_x(tile_next, xb ^ 1, false);
        }
    }
    // region A still holds the previous tile's staged outputs: its bulk stores must have read them before P2 writes Il
    if (tid == ST_TID && out_pending) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        out_pending = false;
    }
    __syncthreads();
    PHASE_MARK(2);
    P2_TRACE(it, 0);

    // ---- P2: projection + loss, one thread per (frame, marker), loop over cameras
    {
        const int f = tid / NL;
        const int l = tid - f * NL;
        const bool live = f < nf;
        const float px = S.p[f][l][0], py = S.p[f][l][1], pz = S.p[f][l][2];
        const float wx = Sx[f][0] + px, wy = Sx[f][1] + py, wz = Sx[f][2] + pz;
        const size_t base = ((size_t)(f0 + f) * C) * NL + l;
        // v = pivot x omega of the 16 pivoted angles, from the components the FK rows stored: one per thread, while the
        // input tiles may still be landing (read by P4, two barriers from here; on one warp inside P3 it was that phase's
        // critical path)
        if (tid < 16 * FT) {
            const int k = tid / FT, ff = tid - k * FT;
            const float* pv = S.p[ff][c_piv.marker[k]];
            float* tt = &S.tau[ff][c_piv.slot[k] * TAU_STRIDE];
            const float q0 = pv[0], q1 = pv[1], q2 = pv[2], o0 = tt[0], o1 = tt[1], o2 = tt[2];
            tt[3] = q1 * o2 - q2 * o1;
            tt[4] = q2 * o0 - q0 * o2;
            tt[5] = q0 * o1 - q1 * o0;
        }
        if (staged) {
            mbar_wait(&S.mbar[1], ph_m & 1);
            ++ph_m;
        }
        // Two cameras (2k, 2k+1) per iteration in the two halves of packed fp32 registers (FFMA2 path);
        // accumulators are pairs too and are folded after the loop.
        const f2 WX = bc(wx), WY = bc(wy), WZ = bc(wz);
        const f2 zero2 = bc(0.f), one2 = bc(1.f);
        f2 A00 = zero2, A01 = zero2, A02 = zero2, A11 = zero2, A12 = zero2, A22 = zero2;
        f2 B0 = zero2, B1 = zero2, B2 = zero2, CST = zero2;
        // (two instantiations: staged tiles read shared memory, the others global memory - no per-iteration branch)
        // NPC > 0: the camera count is 2 * NPC and the loop is unrolled - the camera constants are then immediate offsets
        // into the parameter bank (uniform loads / operands) instead of 28 indexed LDC.64 into vector registers per pair
        auto camera_loop = [&](auto staged_tag, auto np_tag) {
        constexpr bool STAGED = decltype(staged_tag)::value;
        constexpr int NPC = decltype(np_tag)::value;
        const int n_pairs = NPC > 0 ? NPC : (C + 1) / 2;
#pragma unroll(NPC > 0 ? NPC : 1)
        for (int k = 0; k < n_pairs; ++k) {
            const int c = 2 * k;
            const bool has2 = NPC > 0 ? true : c + 1 < C;
            float2 m0 = make_float2(0.f, 0.f), m1 = make_float2(0.f, 0.f);
            float w0 = 0.f, w1 = 0.f;
            if (STAGED) {
                m0 = in_meas[f * ms2 + c * NL + l];
                w0 = in_w[f * ws + c * NL + l];
                if (has2) {
                    m1 = in_meas[f * ms2 + (c + 1) * NL + l];
                    w1 = in_w[f * ws + (c + 1) * NL + l];
                }
            } else if (live) {
                m0 = __ldg(reinterpret_cast<const float2*>(meas) + base + (size_t)c * NL);
                w0 = __ldg(wts + base + (size_t)c * NL);
                if (has2) {
                    m1 = __ldg(reinterpret_cast<const float2*>(meas) + base + (size_t)(c + 1) * NL);
                    w1 = __ldg(wts + base + (size_t)(c + 1) * NL);
                }
            }
            const CamPairF& cp = scene.pair[k];
            const f2 R0 = pk(cp.R[0]), R1 = pk(cp.R[1]), R2 = pk(cp.R[2]), R3 = pk(cp.R[3]), R4 = pk(cp.R[4]),
                     R5 = pk(cp.R[5]), R6 = pk(cp.R[6]), R7 = pk(cp.R[7]), R8 = pk(cp.R[8]);
            const f2 XC = fma2(R0, WX, fma2(R1, WY, fma2(R2, WZ, pk(cp.t[0]))));
            const f2 YC = fma2(R3, WX, fma2(R4, WY, fma2(R5, WZ, pk(cp.t[1]))));
            const f2 ZC = fma2(R6, WX, fma2(R7, WY, fma2(R8, WZ, pk(cp.t[2]))));
            // Kannala-Brandt projection (pt3d_to_2d, all_optimizations.py:193-209) and its Jacobian
            const f2 IZ = rcp2(ZC);
            const f2 Aa = mul2(XC, IZ), Bb = mul2(YC, IZ);
            const f2 RR2 = fma2(Bb, Bb, fma2(Aa, Aa, bc(1e-12f)));
            const f2 IR = rsqrt2(RR2);
            const f2 Rr = mul2(RR2, IR);
            const f2 TH = atan_pos2(Rr, IR);
            const f2 TH2 = mul2(TH, TH);
            const f2 TD = mul2(TH, fma2(TH2, fma2(TH2, fma2(TH2, fma2(TH2, pk(cp.D[3]), pk(cp.D[2])), pk(cp.D[1])), pk(cp.D[0])), one2));
            const f2 DTD = fma2(TH2, fma2(TH2, fma2(TH2, fma2(TH2, pk(cp.D3[3]), pk(cp.D3[2])), pk(cp.D3[1])), pk(cp.D3[0])), one2);
            const f2 SD = mul2(TD, IR);                                                   // s = theta_d / r
            const f2 Q = mul2(sub2(mul2(DTD, rcp2(add2(RR2, one2))), SD), mul2(IR, IR));  // (ds/dr)/r
            const f2 AQ = mul2(Aa, Q), BQ = mul2(Bb, Q);
            const f2 M00 = fma2(Aa, AQ, SD), M01 = mul2(Bb, AQ), M11 = fma2(Bb, BQ, SD);
            const f2 FX = pk(cp.fx), FY = pk(cp.fy);
            // residuals in a centred frame: (fx a s) + (cx - u_meas); zero-weight rows are exactly the
            // constant rho(0) whatever the measurement holds
            f2 RU = fma2(FX, mul2(Aa, SD), sub2(pk(cp.cx), pk(m0.x, m1.x)));
            f2 RV = fma2(FY, mul2(Bb, SD), sub2(pk(cp.cy), pk(m0.y, m1.y)));
            const bool on0 = w0 != 0.f, on1 = w1 != 0.f;
            RU = pk(on0 ? lo(RU) : 0.f, on1 ? hi(RU) : 0.f);
            RV = pk(on0 ? lo(RV) : 0.f, on1 ? hi(RV) : 0.f);
            // world-frame Jacobian rows: J = diag(fx,fy)/z [m] [I | -(a,b)] R = c . G,  G_i = R_i - (a|b) R_2
            const f2 NA_ = sub2(zero2, Aa), NB_ = sub2(zero2, Bb);
            const f2 G00 = fma2(NA_, R6, R0), G01 = fma2(NA_, R7, R1), G02 = fma2(NA_, R8, R2);
            const f2 G10 = fma2(NB_, R6, R3), G11 = fma2(NB_, R7, R4), G12 = fma2(NB_, R8, R5);
            const f2 FXI = mul2(FX, IZ), FYI = mul2(FY, IZ);
            const f2 CU0 = mul2(FXI, M00), CU1 = mul2(FXI, M01), CV0 = mul2(FYI, M01), CV1 = mul2(FYI, M11);
            const f2 JU0 = fma2(CU0, G00, mul2(CU1, G10)), JU1 = fma2(CU0, G01, mul2(CU1, G11)), JU2 = fma2(CU0, G02, mul2(CU1, G12));
            const f2 JV0 = fma2(CV0, G00, mul2(CV1, G10)), JV1 = fma2(CV0, G01, mul2(CV1, G11)), JV2 = fma2(CV0, G02, mul2(CV1, G12));
            // redescending loss of e = |w r| per coordinate
            const f2 Wp = pk(w0, w1);
            const f2 WRU = mul2(Wp, RU), WRV = mul2(Wp, RV);
            const f2 EU = pk(fminf(fmaxf(fabsf(lo(WRU)), 1e-20f), 40.0f), fminf(fmaxf(fabsf(hi(WRU)), 1e-20f), 40.0f));
            const f2 EV = pk(fminf(fmaxf(fabsf(lo(WRV)), 1e-20f), 40.0f), fminf(fmaxf(fabsf(hi(WRV)), 1e-20f), 40.0f));
            f2 rho_u, pr_u, fl_u, rho_v, pr_v, fl_v;
            redescending_fast2(scene.loss, EU, rho_u, pr_u, fl_u);
            redescending_fast2(scene.loss, EV, rho_v, pr_v, fl_v);
            f2 RHO = add2(rho_u, rho_v);
            if (!has2) RHO = pk(lo(RHO), 0.f);        // odd camera count: the padding lane adds nothing
            CST = add2(CST, RHO);
            // d rho / d r = sign(r) w rho'(e) = (rho'(e)/e) w^2 r ; curvature weight max(rho'/e, 1 - sigma_a) w^2
            const f2 W2 = mul2(Wp, Wp);
            const f2 GU = mul2(pr_u, mul2(W2, RU)), GV = mul2(pr_v, mul2(W2, RV));
            B0 = fma2(GU, JU0, fma2(GV, JV0, B0));
            B1 = fma2(GU, JU1, fma2(GV, JV1, B1));
            B2 = fma2(GU, JU2, fma2(GV, JV2, B2));
            if (WANT_H) {
                const f2 HU = mul2(pk(fmaxf(lo(pr_u), lo(fl_u)), fmaxf(hi(pr_u), hi(fl_u))), W2);
                const f2 HV = mul2(pk(fmaxf(lo(pr_v), lo(fl_v)), fmaxf(hi(pr_v), hi(fl_v))), W2);
                const f2 TU0 = mul2(HU, JU0), TU1 = mul2(HU, JU1), TU2 = mul2(HU, JU2);
                const f2 TV0 = mul2(HV, JV0), TV1 = mul2(HV, JV1), TV2 = mul2(HV, JV2);
                A00 = fma2(TU0, JU0, fma2(TV0, JV0, A00));
                A01 = fma2(TU0, JU1, fma2(TV0, JV1, A01));
                A02 = fma2(TU0, JU2, fma2(TV0, JV2, A02));
                A11 = fma2(TU1, JU1, fma2(TV1, JV1, A11));
                A12 = fma2(TU1, JU2, fma2(TV1, JV2, A12));
                A22 = fma2(TU2, JU2, fma2(TV2, JV2, A22));
            }
        }
        };
        if (staged) {
            if (C == 2 * ACINO_UNROLL_PAIRS) camera_loop(std::true_type{}, std::integral_constant<int, ACINO_UNROLL_PAIRS>{});
            else camera_loop(std::true_type{}, std::integral_constant<int, 0>{});
        } else {
            camera_loop(std::false_type{}, std::integral_constant<int, 0>{});
        }
        const float a00 = lo(A00) + hi(A00), a01 = lo(A01) + hi(A01), a02 = lo(A02) + hi(A02);
        const float a11 = lo(A11) + hi(A11), a12 = lo(A12) + hi(A12), a22 = lo(A22) + hi(A22);
        const float b0 = lo(B0) + hi(B0), b1 = lo(B1) + hi(B1), b2 = lo(B2) + hi(B2);
        const float cst = lo(CST) + hi(CST);
        S.costp[f][l] = cst;
        PHASE_MARK(3);
        // spatial inertia of this marker about the head point: B^T A B with B = [-[p]x  I]  (region A: free since the
        // barrier after P1a)
        float o[NSP + 1];
#pragma unroll
        for (int i = 0; i < NSP + 1; ++i) o[i] = 0.f;
        if (WANT_H) {
            // PA = [p]x A  (rows)
            const float q00 = py * a02 - pz * a01, q01 = py * a12 - pz * a11, q02 = py * a22 - pz * a12;
            const float q10 = pz * a00 - px * a02, q11 = pz * a01 - px * a12, q12 = pz * a02 - px * a22;
            const float q20 = px * a01 - py * a00, q21 = px * a11 - py * a01, q22 = px * a12 - py * a02;
            // TL = PA [p]x^T : TL_i0 = -pz q_i1 + py q_i2 ; TL_i1 = pz q_i0 - px q_i2 ; TL_i2 = -py q_i0 + px q_i1
            o[0] = py * q02 - pz * q01;
            o[1] = pz * q00 - px * q02;
            o[2] = px * q01 - py * q00;
            o[3] = pz * q10 - px * q12;
            o[4] = px * q11 - py * q10;
            o[5] = px * q21 - py * q20;
            o[6] = q00; o[7] = q01; o[8] = q02;
            o[9] = q10; o[10] = q11; o[11] = q12;
            o[12] = q20; o[13] = q21; o[14] = q22;
            o[15] = a00; o[16] = a01; o[17] = a02; o[18] = a11; o[19] = a12; o[20] = a22;
        }
        o[21] = py * b2 - pz * b1;
        o[22] = pz * b0 - px * b2;
        o[23] = px * b1 - py * b0;
        o[24] = b0; o[25] = b1; o[26] = b2;
        float4* dst = reinterpret_cast<float4*>(&S.Il[f][l * (NSP + 1)]);
#pragma unroll
        for (int i = (WANT_H ? 0 : 5); i < (NSP + 1) / 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    // ---- P0 + P1a of the NEXT tile, ahead of the barrier: the warps of the camera loop finish hundreds of cycles apart, and
    //      whoever is early spends the wait on the next tile's sin / cos (its state tile was requested during this tile's
    //      FK; sc was last read there) instead of at a barrier of its own at the top of the tile
    if (dynamic && tid == 96) S.next_tile = 2 * (int)gridDim.x + claimed;
    if (tile_next < n_tiles) sincos_tile(tile_next, xb ^ 1);
    __syncthreads();   // the input tiles (region B) are dead from here on
    const int tile_next2 = dynamic ? S.next_tile : tile_next + (int)gridDim.x;
    PHASE_MARK(4);
    P2_TRACE(it, 1);

    // ---- P3: subtree sums up the kinematic tree, one thread per (component pair, frame), packed adds
    {
        constexpr int NKP = (NSP + 1) / 2;       // 14 component pairs
        const int task = tid;
        const int f = task / NKP;
        const int kp = task - f * NKP;
        if (task < FT * NKP) {
            if (WANT_H || kp >= 10) {
                f2 v[NL];
#pragma unroll
                for (int l = 0; l < NL; ++l) v[l] = pk(*reinterpret_cast<const float2*>(&S.Il[f][l * (NSP + 1) + 2 * kp]));
                const f2 s13 = v[19], s12 = add2(v[18], s13);
                const f2 s11 = v[16], s10 = add2(v[15], s11);
                const f2 s5 = v[7], s4 = add2(v[6], s5);
                const f2 s3 = add2(add2(add2(v[5], v[14]), v[17]), add2(add2(s4, s10), s12));
                const f2 s9 = v[13], s8 = add2(v[12], s9);
                const f2 s7 = v[10], s6 = add2(v[9], s7);
                const f2 s2 = add2(add2(add2(v[4], v[8]), v[11]), add2(add2(s3, s6), s8));
                const f2 s1 = add2(v[3], s2);
                const f2 s0 = add2(add2(add2(v[0], v[1]), v[2]), s1);
                float* d = &S.Ij[f][2 * kp];
                constexpr int IS = NSP + 1;
#define ST2(j, v) *reinterpret_cast<float2*>(d + (j) * IS) = make_float2(lo(v), hi(v))
                ST2(0, s0); ST2(1, s1); ST2(2, s2); ST2(3, s3); ST2(4, s4); ST2(5, s5); ST2(6, s6);
                ST2(7, s7); ST2(8, s8); ST2(9, s9); ST2(10, s10); ST2(11, s11); ST2(12, s12); ST2(13, s13);
#undef ST2
            }
        }
    }
    __syncthreads();   // Il is dead from here on; the staged outputs alias it
    PHASE_MARK(5);

    // ---- P4: column beta of H, g[beta].  One thread per (column task, frame).  y = I_subtree(beta) tau_beta stays in registers
    //      (see ColEntry).  Threads 0 .. 16 FT - 1: the 16 columns with non-trunk ancestors (tasks 0..15 of the table, four
    //      per warp, sorted by list length).  The last 4 FT threads: the six trunk columns (v = 0, every ancestor a trunk slot:
    //      half the arithmetic, no ancestor list) - four, then two beside the translation block.
    auto column_task = [&](const int q, const int f) {
        float* g = S.o.g[f];
        float* H = S.o.H[f];
        const uint4* ce = reinterpret_cast<const uint4*>(&S.col[q]);
        const uint4 e0 = ce[0];            // tau offset, Ij offset, list length, slot
        const int be = e0.w, n_anc = e0.z;
        const float4* I4 = reinterpret_cast<const float4*>(&S.Ij[f][e0.y]);
        const float4 i0 = I4[0], i1 = I4[1], i2 = I4[2], i3 = I4[3], i4 = I4[4], i5 = I4[5], i6 = I4[6];
        const float I[28] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w, i3.x, i3.y,
                             i3.z, i3.w, i4.x, i4.y, i4.z, i4.w, i5.x, i5.y, i5.z, i5.w, i6.x, i6.y, i6.z, i6.w};
        const float* tau_f = &S.tau[f][0];
        const float4 t0 = *reinterpret_cast<const float4*>(tau_f + e0.x);
        const float2 t1 = *reinterpret_cast<const float2*>(tau_f + e0.x + 4);
        const float o0 = t0.x, o1 = t0.y, o2 = t0.z, v0 = t0.w, v1 = t1.x, v2 = t1.y;
        const float gb = o0 * I[21] + o1 * I[22] + o2 * I[23] + v0 * I[24] + v1 * I[25] + v2 * I[26];
        if (!WANT_H) {
            g[3 + be] = gb;
            return;
        }
        // y = I tau_beta ; I = [[TL, PA],[PA^T, A]]
        const float yt0 = I[0] * o0 + I[1] * o1 + I[2] * o2 + I[6] * v0 + I[7] * v1 + I[8] * v2;
        const float yt1 = I[1] * o0 + I[3] * o1 + I[4] * o2 + I[9] * v0 + I[10] * v1 + I[11] * v2;
        const float yt2 = I[2] * o0 + I[4] * o1 + I[5] * o2 + I[12] * v0 + I[13] * v1 + I[14] * v2;
        const float yb0 = I[6] * o0 + I[9] * o1 + I[12] * o2 + I[15] * v0 + I[16] * v1 + I[17] * v2;
        const float yb1 = I[7] * o0 + I[10] * o1 + I[13] * o2 + I[16] * v0 + I[18] * v1 + I[19] * v2;
        const float yb2 = I[8] * o0 + I[11] * o1 + I[14] * o2 + I[17] * v0 + I[19] * v1 + I[20] * v2;
        // trunk rows: omega_al . y_top (the loads are the same for the four columns of a frame in this warp)
        float ht[N_TRUNK];
#pragma unroll
        for (int k = 0; k < N_TRUNK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(tau_f + trunk_slot(k) * TAU_STRIDE);
            ht[k] = a0.x * yt0 + a0.y * yt1 + a0.z * yt2;
        }
        // the other ancestors (and beta itself): full 6-term products
        const uint4 a0_ = ce[3], a1_ = ce[4];           // anc_off[0..7]
        const unsigned ao[MAX_ANC] = {a0_.x, a0_.y, a0_.z, a0_.w, a1_.x, a1_.y, a1_.z, a1_.w};
        // in batches with the loads of a batch in flight together (one warp-wide maximum of the list lengths decides which
        // batches run: a loop with a vote and a break per ancestor ran them one load latency after the other; executing all 8
        // for every column measured 4 % slower).  Unused list entries point at slot 0: a valid address, the product is not stored
        float ha[MAX_ANC];
        const int n_max = (int)__reduce_max_sync(__activemask(), (unsigned)n_anc);
        auto anc_batch = [&](auto b_tag, auto n_tag) {
            constexpr int B0 = decltype(b_tag)::value, NB = decltype(n_tag)::value;
            float4 a0[NB];
            float2 a1[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                a0[k] = *reinterpret_cast<const float4*>(tau_f + ao[B0 + k]);
                a1[k] = *reinterpret_cast<const float2*>(tau_f + ao[B0 + k] + 4);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
                ha[B0 + k] = a0[k].x * yt0 + a0[k].y * yt1 + a0[k].z * yt2 + a0[k].w * yb0 + a1[k].x * yb1 + a1[k].y * yb2;
        };
        using std::integral_constant;
        // the four warps of the generic columns have list lengths up to 2 / 4 / 6 / 8 (sorted table): 2 or 4 per batch
        if (n_max <= 2) {
            anc_batch(integral_constant<int, 0>{}, integral_constant<int, 2>{});
        } else {
            anc_batch(integral_constant<int, 0>{}, integral_constant<int, 4>{});
            if (n_max > 6) anc_batch(integral_constant<int, 4>{}, integral_constant<int, 4>{});
            else if (n_max > 4) anc_batch(integral_constant<int, 4>{}, integral_constant<int, 2>{});
        }
        // stores last: no store -> load ordering inside the task
        g[3 + be] = gb;
        const int sb = 3 + be;  // active slot of beta: rows 0..2 (translation) of column sb
        H[sb] = yb0;
        H[NA + sb - 1] = yb1;
        H[2 * NA + sb - 3] = yb2;
        const uint4 t0_ = ce[1];
        const uint2 t1_ = *reinterpret_cast<const uint2*>(&ce[2]);
        const unsigned ti[N_TRUNK] = {t0_.x, t0_.y, t0_.z, t0_.w, t1_.x, t1_.y};
#pragma unroll
        for (int k = 0; k < N_TRUNK; ++k) H[ti[k]] = ht[k];      // (every trunk slot is an ancestor of a generic column)
        const uint4 i0_ = ce[5], i1_ = ce[6];           // anc_idx[0..7]
        const unsigned ai[MAX_ANC] = {i0_.x, i0_.y, i0_.z, i0_.w, i1_.x, i1_.y, i1_.z, i1_.w};
#pragma unroll
        for (int k = 0; k < MAX_ANC; ++k)
            if (k < n_anc) H[ai[k]] = ha[k];
    };
    // a trunk column: tau_beta = (omega, 0), so y = I[:, 0:3] omega (18 products instead of 36), the wrench part of g needs
    // components 21..23 only, and the ancestors-or-self are all among the six trunk slots (the table's trunk_idx)
    auto trunk_task = [&](const int q, const int f) {
        float* g = S.o.g[f];
        float* H = S.o.H[f];
        const uint4* ce = reinterpret_cast<const uint4*>(&S.col[q]);
        const uint4 e0 = ce[0];
        const int be = e0.w;
        const float4* I4 = reinterpret_cast<const float4*>(&S.Ij[f][e0.y]);
        const float* tau_f = &S.tau[f][0];
        const float4 t0 = *reinterpret_cast<const float4*>(tau_f + e0.x);
        const float o0 = t0.x, o1 = t0.y, o2 = t0.z;
        const float4 i5 = I4[5];           // components 20..23
        const float gb = o0 * i5.y + o1 * i5.z + o2 * i5.w;
        g[3 + be] = gb;
        if (!WANT_H) return;
        const float4 i0 = I4[0], i1 = I4[1], i2 = I4[2], i3 = I4[3];
        const float I[16] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w, i3.x, i3.y, i3.z, i3.w};
        const float yt0 = I[0] * o0 + I[1] * o1 + I[2] * o2;
        const float yt1 = I[1] * o0 + I[3] * o1 + I[4] * o2;
        const float yt2 = I[2] * o0 + I[4] * o1 + I[5] * o2;
        const float yb0 = I[6] * o0 + I[9] * o1 + I[12] * o2;
        const float yb1 = I[7] * o0 + I[10] * o1 + I[13] * o2;
        const float yb2 = I[8] * o0 + I[11] * o1 + I[14] * o2;
        float ht[N_TRUNK];
#pragma unroll
        for (int k = 0; k < N_TRUNK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(tau_f + trunk_slot(k) * TAU_STRIDE);
            ht[k] = a0.x * yt0 + a0.y * yt1 + a0.z * yt2;
        }
        const int sb = 3 + be;
        H[sb] = yb0;
        H[NA + sb - 1] = yb1;
        H[2 * NA + sb - 3] = yb2;
        const uint4 t0_ = ce[1];
        const uint2 t1_ = *reinterpret_cast<const uint2*>(&ce[2]);
        const unsigned ti[N_TRUNK] = {t0_.x, t0_.y, t0_.z, t0_.w, t1_.x, t1_.y};
#pragma unroll
        for (int k = 0; k < N_TRUNK; ++k)
            if (ti[k] != NO_ENTRY) H[ti[k]] = ht[k];
    };
    // the translation block, the translation part of g and the frame's cost
    auto translation_task = [&](const int f) {
        float* g = S.o.g[f];
        float* H = S.o.H[f];
        float c = 0.f;
        static_assert(NL % 4 == 0, "cost partials are read four at a time");
#pragma unroll
        for (int l = 0; l < NL; l += 4) {
            const float4 v = *reinterpret_cast<const float4*>(&S.costp[f][l]);
            c += v.x;
            c += v.y;
            c += v.z;
            c += v.w;
        }
        const float* I0 = S.Ij[f];
        const float g0 = I0[24], g1 = I0[25], g2 = I0[26];
        const float h0 = I0[15], h1 = I0[16], h2 = I0[17], h3 = I0[18], h4 = I0[19], h5 = I0[20];
        S.o.cost[f] = c;
        g[0] = g0; g[1] = g1; g[2] = g2;
        if (WANT_H) {
            H[upper_index(0, 0)] = h0; H[upper_index(0, 1)] = h1; H[upper_index(0, 2)] = h2;
            H[upper_index(1, 1)] = h3; H[upper_index(1, 2)] = h4; H[upper_index(2, 2)] = h5;
        }
    };
    if (tid < N_GENERIC * FT) {
        column_task(tid / FT, tid % FT);
    } else {
        const int u = tid - N_GENERIC * FT;          // 0 .. 4 FT - 1
        trunk_task(N_GENERIC + u / FT, u % FT);
        if (u < 2 * FT) trunk_task(N_GENERIC + 4 + u / FT, u % FT);
        else if (u < 3 * FT) translation_task(u - 2 * FT);
    }
    // structural zeros of H: N_ZERO entries per frame, no arithmetic
    if (WANT_H) {
        // (unrolled: the list index of pass i is (tid / FT) + i * NL, an immediate offset; the rolled loop spent 9 instructions
        // per entry on its own bookkeeping)
        static_assert(NT % FT == 0, "a thread keeps its frame across the passes");
        const int zf = tid % FT, zk = tid / FT;
        float* const Hz = S.o.H[zf];
#pragma unroll
        for (int i = 0; i < (N_ZERO * FT + NT - 1) / NT; ++i)
            if ((i + 1) * NT <= N_ZERO * FT || zk + i * (NT / FT) < N_ZERO) Hz[S.zero_idx[zk + i * (NT / FT)]] = 0.f;
    }
    __syncthreads();   // Ij (region B) is dead from here on
    PHASE_MARK(6);
    // ---- the next tile's measurements and weights: region B is free; the copies land while P5 and the next tile's FK run
    if (bulk_in && tid == MW_TID) {        // (thread ST_TID issues the bulk stores of P5 meanwhile)
        const int tn = tile_next;
        if (tn < n_tiles && (tn + 1) * FT <= n_frames) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_mw(tn);
        }
    }
    PHASE_MARK(7);

    // ---- P5: write-out.  The staged blocks have the global layout.  Full tiles with 16-byte aligned outputs:
    //      three bulk async stores (TMA) issued by one thread (they drain while the next tile starts); otherwise
    //      straight copies
    if ((use_bulk & 2) && nf == FT) {
        if (tid == ST_TID) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (cost_out) bulk_s2g(cost_out + f0, &S.o.cost[0], FT * 4);
            if (g_out) bulk_s2g(g_out + (size_t)f0 * NA, &S.o.g[0][0], FT * NA * 4);
            if (WANT_H && H_out) bulk_s2g(H_out + (size_t)f0 * NU, &S.o.H[0][0], FT * NU * 4);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            out_pending = true;        // waited for before the next tile overwrites the stage (or at exit)
        }
    } else {
        // partial tile or outputs that are not 16-byte aligned: scalar copies
        if (cost_out)
            for (int f = tid; f < nf; f += NT) cost_out[f0 + f] = S.o.cost[f];
        if (g_out) {
            float* dst = g_out + (size_t)f0 * NA;
            const float* src = &S.o.g[0][0];
            for (int i = tid; i < nf * NA; i += NT) dst[i] = src[i];
        }
        if (WANT_H && H_out) {
            float* dst = H_out + (size_t)f0 * NU;
            const float* src = &S.o.H[0][0];
            for (int i = tid; i < nf * NU; i += NT) dst[i] = src[i];
        }
    }
    PHASE_MARK(8);
    tile = tile_next;
    tile_next = tile_next2;
    }   // tiles
    if (tid == ST_TID && out_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (dynamic && tid == 0) {             // every ticket of this CTA has been drawn: the last CTA out re-arms the counters
        if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {
            sched[0] = 0;
            sched[1] = 0;
        }
    }
}

// persistent: one wave of MINB CTAs per SM, each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int FT, bool WANT_H, int MINB>
__global__ void __launch_bounds__(FT * NL, MINB)
fte_eval_kernel(const __grid_constant__ SceneF scene, const int n_frames, const int use_bulk,
                const float* __restrict__ xg, const float* __restrict__ meas, const float* __restrict__ wts,
                float* __restrict__ cost_out, float* __restrict__ g_out, float* __restrict__ H_out, int* __restrict__ sched) {
    fte_eval_body<FT, WANT_H>(scene, n_frames, use_bulk, xg, meas, wts, cost_out, g_out, H_out, sched);
}

#ifdef ACINO_PHASE_TIMING
extern "C" void acino_debug_phase_cycles(long long* out16) { cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(long long) * 16); }
extern "C" void acino_debug_p2_trace(long long* trace, int* smid) {
    cudaMemcpyFromSymbol(trace, g_p2_trace, sizeof(long long) * 1024 * 128 * 2);
    cudaMemcpyFromSymbol(smid, g_p2_smid, sizeof(int) * 1024);
}
extern "C" void acino_debug_p2_trace_reset() {
    void* p = nullptr;
    cudaGetSymbolAddress(&p, g_p2_trace);
    cudaMemset(p, 0, sizeof(long long) * 1024 * 128 * 2);
}
extern "C" void acino_debug_phase_reset() { long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
#endif

// ------------------------------------------------------------------------------------------
// fk_project: pose_to_3d + project_points_fisheye for every camera (reprojection only).
constexpr int FTP = 8;
__global__ void __launch_bounds__(FTP * NL)
fk_project_kernel(const __grid_constant__ SceneF scene, const int n_frames, const float* __restrict__ xg,
                  float* __restrict__ pos_out, float* __restrict__ uv_out) {
    constexpr int NT = FTP * NL;
    __shared__ float sx[FTP][NA];
    __shared__ float2 ssc[FTP][NANG];
    __shared__ float sp[FTP][NL][3];
    __shared__ __align__(16) float stau[FTP][TAUF];
    const int tid = threadIdx.x;
    const int f0 = blockIdx.x * FTP;
    const int nf = min(FTP, n_frames - f0);
    const int C = scene.n_cams;
    for (int i = tid; i < FTP * NA; i += NT) {
        const int f = i / NA;
        (&sx[0][0])[i] = (f < nf) ? xg[(size_t)f0 * NA + i] : 0.f;
    }
    __syncthreads();
    for (int t = tid; t < FTP * NANG; t += NT) {
        const int a = t / FTP, f = t - a * FTP;
        float sn, cs;
        sincosf(sx[f][3 + a], &sn, &cs);
        ssc[f][a] = make_float2(sn, cs);
    }
    __syncthreads();
    if (tid < FTP) {
        FkWriter w{&sp[tid][0][0], &stau[tid][0]};
        cheetah_fk(ssc[tid], w);
    }
    __syncthreads();
    const int f = tid / NL;
    const int l = tid - f * NL;
    if (f >= nf) return;
    const float wx = sx[f][0] + sp[f][l][0], wy = sx[f][1] + sp[f][l][1], wz = sx[f][2] + sp[f][l][2];
    if (pos_out) {
        float* o = pos_out + ((size_t)(f0 + f) * NL + l) * 3;
        o[0] = wx; o[1] = wy; o[2] = wz;
    }
    if (uv_out) {
        for (int c = 0; c < C; ++c) {
            const CamF& cam = scene.cam[c];
            const float xc = fmaf(cam.R[0], wx, fmaf(cam.R[1], wy, fmaf(cam.R[2], wz, cam.t[0])));
            const float yc = fmaf(cam.R[3], wx, fmaf(cam.R[4], wy, fmaf(cam.R[5], wz, cam.t[1])));
            const float zc = fmaf(cam.R[6], wx, fmaf(cam.R[7], wy, fmaf(cam.R[8], wz, cam.t[2])));
            ProjOut<float> pr;
            fisheye_cam<float, false>(xc, yc, zc, cam.fx, cam.fy, cam.D[0], cam.D[1], cam.D[2], cam.D[3], pr);
            float2 o = {pr.u + cam.cx, pr.v + cam.cy};
            reinterpret_cast<float2*>(uv_out)[((size_t)(f0 + f) * C + c) * NL + l] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
using FteKernel = void (*)(const SceneF, int, int, const float*, const float*, const float*, float*, float*, float*, int*);

// per-device launch configuration (cudaFuncSetAttribute is per device; a process may drive several GPUs)
struct FteDeviceCfg {
    FteKernel configured[8];
    int n_configured;
    int n_sm;
};
static FteDeviceCfg g_dev_cfg[64];

constexpr int FTE_FT = 8;        // frames per tile: 160 threads
constexpr int FTE_CTAS = 4;      // resident CTAs per SM (96 registers, no spills; 47.9 KB of shared memory each)

template <int FT, int MINB>
static cudaError_t launch_fte_eval_k(const SceneF& scene, int n_frames, const float* x, const float* meas, const float* w,
                                     float* cost, float* g, float* H, cudaStream_t stream, int ctas_per_sm, size_t smem_pad,
                                     int* sched) {
    // bulk (TMA) staging needs 16-byte aligned tiles: frame tiles are multiples of 16 bytes, so it is the
    // base pointers that decide
    // bit 0: inputs, bit 1: outputs (cost tiles are FT * 4 = 32 bytes, g / H tiles multiples of 16 bytes)
    const int use_bulk = (((((uintptr_t)x | (uintptr_t)meas | (uintptr_t)w) & 15u) == 0) ? 1 : 0) |
                         (((((uintptr_t)cost | (uintptr_t)g | (uintptr_t)H) & 15u) == 0 && FT % 4 == 0) ? 2 : 0);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    FteDeviceCfg& cfg = g_dev_cfg[dev];
    if (!cfg.n_sm) cudaDeviceGetAttribute(&cfg.n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = (n_frames + FT - 1) / FT;
    const int grid = n_tiles < cfg.n_sm * ctas_per_sm ? n_tiles : cfg.n_sm * ctas_per_sm;
    const size_t smem = sizeof(Smem<FT>) + smem_pad;
    FteKernel k = H ? (FteKernel)fte_eval_kernel<FT, true, MINB> : (FteKernel)fte_eval_kernel<FT, false, MINB>;
    bool done = false;
    for (int i = 0; i < cfg.n_configured; ++i) done |= cfg.configured[i] == k;
    if (!done || smem_pad) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (!done && cfg.n_configured < 8) cfg.configured[cfg.n_configured++] = k;
    }
    k<<<grid, FT * NL, smem, stream>>>(scene, n_frames, use_bulk, x, meas, w, cost, g, H, sched);
    return cudaGetLastError();
}

// sched: device pointer to a zero-initialised {ticket, finished} counter pair owned by the caller (one per launch in flight;
// the kernel leaves it zeroed), or nullptr for the static tile schedule
cudaError_t launch_fte_eval(const SceneF& scene, int n_frames, const float* x, const float* meas,
                            const float* w, float* cost, float* g, float* H, cudaStream_t stream, int* sched) {
    if (n_frames <= 0) return cudaSuccess;
#ifdef ACINO_EXPERIMENTS
    // A/B knobs of scripts/bench_variants.sh / bench_residency.sh (build with -DACINO_EXPERIMENTS):
    //   ACINO_FTE_VARIANT 1: 5 CTAs/SM (<= 80 registers);  ACINO_FTE_CTAS n: grid of n CTAs per SM;
    //   ACINO_FTE_SMEM_PAD bytes: extra dynamic shared memory per CTA (forces the residency down)
    static int variant = -1, ctas = 0;
    static size_t pad = 0;
    if (variant < 0) {
        const char* e = getenv("ACINO_FTE_VARIANT");
        variant = e ? atoi(e) : 0;
        e = getenv("ACINO_FTE_CTAS");
        ctas = e ? atoi(e) : 0;
        e = getenv("ACINO_FTE_SMEM_PAD");
        pad = e ? (size_t)atol(e) : 0;
    }
    if (variant == 1) return launch_fte_eval_k<FTE_FT, 5>(scene, n_frames, x, meas, w, cost, g, H, stream, ctas ? ctas : 5, pad, sched);
    if (variant == 3) return launch_fte_eval_k<FTE_FT, 3>(scene, n_frames, x, meas, w, cost, g, H, stream, ctas ? ctas : 3, pad, sched);
    return launch_fte_eval_k<FTE_FT, FTE_CTAS>(scene, n_frames, x, meas, w, cost, g, H, stream, ctas ? ctas : FTE_CTAS, pad, sched);
#else
    return launch_fte_eval_k<FTE_FT, FTE_CTAS>(scene, n_frames, x, meas, w, cost, g, H, stream, FTE_CTAS, 0, sched);
#endif
}

const char* fte_eval_kernel_name(int) { return "fte_eval_kernel<8, 1, 4>"; }

cudaError_t launch_fk_project(const SceneF& scene, int n_frames, const float* x, float* pos, float* uv,
                              cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int grid = (n_frames + FTP - 1) / FTP;
    fk_project_kernel<<<grid, FTP * NL, 0, stream>>>(scene, n_frames, x, pos, uv);
    return cudaGetLastError();
}

}  // namespace acino
