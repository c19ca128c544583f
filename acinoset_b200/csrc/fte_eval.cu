// fte_eval: per-frame reprojection cost, gradient and Gauss-Newton block of the FTE objective.
//
// Replaces (reference, /root/reference/src/all_optimizations.py): the SymPy FK chain :66-179,
// pose_constraint :359-365, measurement_constraints :394-399 through pt3d_to_2d :193-209, the
// measurement term of obj :494-497 with redescending_loss (build.py:382-395), and the
// automatic differentiation IPOPT's ASL performs on those expression trees each iteration.
//
// Layout / mapping (B200, CUDA cores only - the contraction is tiny, irregular and sparse):
//   one CTA = FT frames.  Phases, separated by __syncthreads():
//   P1  FK            thread <-> frame           22 sincos, rotation chain in registers,
//                                                writes marker positions (relative to the head)
//                                                and one twist (omega, pivot x omega) per angle
//   P2  projection    thread <-> (frame, marker) loops over cameras: fisheye projection, 2x3
//                                                Jacobian, redescending loss; accumulates the
//                                                marker's 3x3 normal block A_l, 3-vector b_l and
//                                                cost in registers (no cross-thread reduction),
//                                                then forms the marker's 6x6 "spatial inertia"
//                                                B_l^T A_l B_l and wrench B_l^T b_l, B_l=[-[p]x I]
//   P3  subtree sums  thread <-> (frame, comp)   composite-rigid-body style accumulation of the
//                                                27 components up the kinematic tree (fixed order)
//   P4  blocks        thread <-> (frame, angle)  y = I_subtree tau_beta; H[a][b] = tau_a . y for
//                                                every ancestor angle a; g[b] = tau_b . wrench
//   P5  write-out     coalesced copy of {cost, g[25], H[325]} from shared memory
// d p_l / d angle = omega x (p_l - pivot) = B_l tau (twist about the head point), so
// H = sum_l J_l^T A_l J_l collapses to tau_a^T I_{deeper subtree} tau_b: ~4 kFMA instead of
// ~8 kFMA per frame for the chain rule and no 60x25 Jacobian is ever materialised.
#include <cstdlib>
#include <type_traits>

#include "acino_common.cuh"
#include "cheetah_fk.cuh"

namespace acino {

constexpr int PREFETCH_DIST = 148 * 4;   // tiles one resident wave ahead (experiment, ACINO_FTE_EXP bit 0)
constexpr int N_PAIR = NANG * (NANG + 1) / 2;  // 253 unordered angle pairs (incl. the diagonal)

// One entry per unordered pair of angle slots.  Related pairs (one joint is an ancestor-or-self of
// the other): H = tau_al . y_be with `be` the deeper angle.  Entry = float offsets into the frame's
// tau / y arrays and the packed H index: bits 0-7 al*8, 8-15 be*8, 16-24 index.  Unrelated pairs
// (disjoint subtrees) are structural zeros, sorted last: entries N_REL.. only carry the H index and are
// written as 0 without any arithmetic.
struct PairTable {
    unsigned e[N_PAIR + 3];     // padded to 256 entries = 1024 bytes: one bulk copy
    int joint[NANG];
};
constexpr PairTable make_pair_table() {
    PairTable t{};
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int be = 0; be < NANG; ++be)
            for (int al = 0; al < NANG; ++al) {
                const int ja = k_angle_joint[al], jb = k_angle_joint[be];
                const bool a_anc_b = joint_is_anc(ja, jb), b_anc_a = joint_is_anc(jb, ja);
                const int sa = 3 + al, sb = 3 + be;
                const int lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
                const unsigned idx = (unsigned)(lo * NA - (lo * (lo - 1)) / 2 + (hi - lo));
                if (pass == 0 && a_anc_b && (ja != jb || al <= be))
                    t.e[n++] = (unsigned)(al * TAU_STRIDE) | ((unsigned)(be * TAU_STRIDE) << 8) | (idx << 16);
                if (pass == 1 && !a_anc_b && !b_anc_a && al < be)
                    t.e[n++] = (unsigned)(NANG * TAU_STRIDE) | ((unsigned)(be * TAU_STRIDE) << 8) | (idx << 16);
            }
    for (int a = 0; a < NANG; ++a) t.joint[a] = k_angle_joint[a];
    return t;
}
constexpr int count_related_pairs() {
    int n = 0;
    for (int be = 0; be < NANG; ++be)
        for (int al = 0; al < NANG; ++al) {
            const int ja = k_angle_joint[al], jb = k_angle_joint[be];
            if (joint_is_anc(ja, jb) && (ja != jb || al <= be)) ++n;
        }
    return n;
}
constexpr int N_REL = count_related_pairs();   // 185 related pairs; the other 68 are structural zeros
static_assert(N_REL == 185, "kinematic tree changed: check the pair table");
__constant__ PairTable c_tab = make_pair_table();
__device__ __align__(16) const PairTable d_tab = make_pair_table();     // global-memory copy: source of the bulk (TMA) copy

// bulk-copied input tiles.  One bulk copy per frame into padded frame slots: the 20 lanes of frame f+1 continue in
// the banks where frame f stopped (stride = 8 (mod 32) words for the float2 tile, 20 (mod 32) for the weights), so a
// warp that straddles two frames reads conflict-free
template <int FT, int MAXC>
struct InTiles {
    float2 meas[FT * (MAXC * NL + 16)];   // [FT][C][NL] (u,v)
    float w[FT * (MAXC * NL + 32)];       // [FT][C][NL]
};
struct NoTiles {};
constexpr int MAXC_PERSIST = 6;           // cameras the separate (prefetchable) input buffer of the persistent kernel holds

template <int FT, bool PERSIST>
struct __align__(16) Smem {
    float x[2][FT][NA];                // state (double buffered: the persistent kernel prefetches the next tile)
    float p[FT][NL][3];                // marker positions relative to the head point
    float2 sc[FT][NANG];               // (sin, cos) of every angle
    float costp[FT][NL];               // per-(frame, marker) cost partials
    unsigned tab[N_PAIR + 3];          // pair table (copy of c_tab.e)
    // subtree spatial inertia + wrench per joint (27 padded to 28: float4 loads); frame stride 396 words = 12 (mod 32):
    // the 8 frames of a quarter-warp hit 8 distinct 4-bank groups (392 gave 2-way conflicts on every LDS.128 of P4a)
    __align__(16) float Ij[FT][NJ * (NSP + 1) + 4];
    __align__(16) float tau[FT][TAUF]; // (omega, v = pivot x omega) per angle, stride 8
    unsigned long long mbar[2];        // [0] state tile landed, [1] measurement + weight tiles landed
    // persistent kernel: input tiles in their own buffer, refilled for the NEXT tile while P2b..P5 of the current one run
    __align__(16) typename std::conditional<PERSIST, InTiles<FT, MAXC_PERSIST>, NoTiles>::type in_s;
    union {
        __align__(16) float Il[FT * NL][NSP + 1];   // per-marker spatial inertia + wrench, 27 padded to 28: STS.128 / LDS.64  (P2 -> P3)
        InTiles<FT, ACINO_MAX_CAMS / 2> in_u;       // one-tile-per-CTA kernel: input tiles alias Il (kernel start -> end of the camera loop)
        struct {                       // staged outputs in their global layout  (P4 -> P5)
            float H[FT][NU];
            float g[FT][NA];
            float cost[FT];
            __align__(16) float y[FT][TAUF];   // y_beta = I_subtree tau_beta, stride 8
        } o;
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
#else
#define PHASE_MARK(i) do { } while (0)
#endif

template <int FT, bool WANT_H, int NPAIR, bool PERSIST>
__device__ __forceinline__ void
fte_eval_body(const SceneF& scene, const int n_frames, const int use_bulk,
              const float* __restrict__ xg, const float* __restrict__ meas,
              const float* __restrict__ wts, float* __restrict__ cost_out,
              float* __restrict__ g_out, float* __restrict__ H_out) {
    constexpr int NT = FT * NL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<FT, PERSIST>& S = *reinterpret_cast<Smem<FT, PERSIST>*>(smem_raw);
    const int tid = threadIdx.x;
    const int C = scene.n_cams;
    const int n_tiles = (n_frames + FT - 1) / FT;
    // the input tiles of a full tile are contiguous, 16-byte aligned blocks of global memory: stage them with 1-D
    // bulk async copies (TMA) that overlap the forward kinematics (and, in the persistent kernel, the previous tile)
    const bool bulk_in = (use_bulk & 1) && C <= (PERSIST ? MAXC_PERSIST : ACINO_MAX_CAMS / 2);
    float2* in_meas;
    float* in_w;
    if constexpr (PERSIST) {
        in_meas = S.in_s.meas;
        in_w = S.in_s.w;
    } else {
        in_meas = S.in_u.meas;
        in_w = S.in_u.w;
    }
    // padded frame strides of the staged tiles (see InTiles): float2 units / float units
    const int ms2 = (C * NL * 2 + ((8 - C * NL * 2) & 31)) >> 1;
    const int ws = C * NL + ((20 - C * NL) & 31);
    const unsigned bx = FT * NA * 4, bm = C * NL * 8, bw = C * NL * 4;     // bytes: state tile; per-frame meas / weight rows
    auto issue_x = [&](const int t, const int buf, const bool with_tab) {
        mbar_expect_tx(&S.mbar[0], bx + (with_tab ? (N_PAIR + 3) * 4 : 0));
        bulk_g2s(&S.x[buf][0][0], xg + (size_t)t * FT * NA, bx, &S.mbar[0]);
        if (with_tab) bulk_g2s(&S.tab[0], &d_tab.e[0], (N_PAIR + 3) * 4, &S.mbar[0]);
    };
    auto issue_mw = [&](const int t) {
        mbar_expect_tx(&S.mbar[1], FT * (bm + bw));
        for (int ff = 0; ff < FT; ++ff) {
            bulk_g2s(&in_meas[ff * ms2], meas + (size_t)(t * FT + ff) * C * NL * 2, bm, &S.mbar[1]);
            bulk_g2s(&in_w[ff * ws], wts + (size_t)(t * FT + ff) * C * NL, bw, &S.mbar[1]);
        }
    };
#ifdef ACINO_PHASE_TIMING
    long long _tprev = clock64();
#endif

    // ---- prologue: barriers, zero twist slot, the first tile's copies
    int tile = blockIdx.x;
    if (bulk_in && tid == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < n_tiles && (tile + 1) * FT <= n_frames) {
            issue_x(tile, 0, true);
            issue_mw(tile);
        }
        // experiment (use_bulk bit 2): pull the inputs of the tile that will run in this CTA slot one wave later into
        // L2, so that its own bulk copies do not pay the full HBM latency at CTA start
        if (!PERSIST && (use_bulk & 4)) {
            const int tp = tile + PREFETCH_DIST;
            if ((tp + 1) * FT <= n_frames) {
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xg + (size_t)tp * FT * NA), "r"(bx) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(meas + (size_t)tp * FT * C * NL * 2), "r"(FT * bm) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wts + (size_t)tp * FT * C * NL), "r"(FT * bw) : "memory");
            }
        }
    }
    if (tid < FT * TAU_STRIDE) S.tau[tid / TAU_STRIDE][NANG * TAU_STRIDE + (tid % TAU_STRIDE)] = 0.f;
    bool tab_ready = false;
    unsigned ph_x = 0, ph_m = 0;       // completed phases of mbar[0] / mbar[1]
    bool out_pending = false;          // (thread 0) bulk stores of the previous tile may still be reading `o`
    __syncthreads();

    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int f0 = tile * FT;
    const int nf = min(FT, n_frames - f0);
    const bool staged = bulk_in && nf == FT;
    const int xb = PERSIST ? (it & 1) : 0;
    float (*Sx)[NA] = S.x[xb];
    // ---- P0: the tile's state (and, first time, the pair table)
    if (staged) {
        mbar_wait(&S.mbar[0], ph_x & 1);
        ++ph_x;
        tab_ready = true;
    } else {
        for (int i = tid; i < FT * NA; i += NT) {
            const int f = i / NA;
            (&Sx[0][0])[i] = (f < nf) ? xg[(size_t)f0 * NA + i] : 0.f;
        }
        if (!tab_ready)
            for (int i = tid; i < N_PAIR; i += NT) S.tab[i] = c_tab.e[i];
        tab_ready = true;
        __syncthreads();
    }
    PHASE_MARK(0);

    // ---- P1a: sin/cos of the 22 angles, one thread per (angle, frame)
    for (int t = tid; t < FT * NANG; t += NT) {
        const int a = t / FT, f = t - a * FT;
        float sn, cs;
        if (use_bulk & 8) {            // experiment: range-reduced MUFU sin / cos
            const float xa = Sx[f][3 + a];
            const float xr = fmaf(-6.283185307179586f, rintf(xa * 0.15915494309189535f), xa);
            sn = __sinf(xr);
            cs = __cosf(xr);
        } else {
            sincosf(Sx[f][3 + a], &sn, &cs);
        }
        S.sc[f][a] = make_float2(sn, cs);
    }
    __syncthreads();
    if (PERSIST && bulk_in && tid == 0) {          // every thread is past its wait on mbar[0]: re-arm it for the next tile
        const int tn = tile + gridDim.x;
        if (tn < n_tiles && (tn + 1) * FT <= n_frames) issue_x(tn, xb ^ 1, false);
    }
    PHASE_MARK(1);

    // ---- P1b: rotation chain, one thread per frame
    if (tid < FT) {
        FkWriter w{&S.p[tid][0][0], &S.tau[tid][0]};
        cheetah_fk(S.sc[tid], w);
    }
    __syncthreads();
    PHASE_MARK(2);

    // ---- P2: projection + loss, one thread per (frame, marker), loop over cameras
    {
        const int f = tid / NL;
        const int l = tid - f * NL;
        const bool live = f < nf;
        const float px = S.p[f][l][0], py = S.p[f][l][1], pz = S.p[f][l][2];
        const float wx = Sx[f][0] + px, wy = Sx[f][1] + py, wz = Sx[f][2] + pz;
        const size_t base = ((size_t)(f0 + f) * C) * NL + l;
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
#pragma unroll
        for (int c = 0; c < (NPAIR ? 2 * NPAIR : C); c += 2) {
            const bool has2 = NPAIR ? true : (c + 1 < C);
            float2 m0 = make_float2(0.f, 0.f), m1 = make_float2(0.f, 0.f);
            float w0 = 0.f, w1 = 0.f;
            if (staged) {
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
            const CamPairF& cp = scene.pair[c >> 1];
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
            const f2 NA = sub2(zero2, Aa), NB = sub2(zero2, Bb);
            const f2 G00 = fma2(NA, R6, R0), G01 = fma2(NA, R7, R1), G02 = fma2(NA, R8, R2);
            const f2 G10 = fma2(NB, R6, R3), G11 = fma2(NB, R7, R4), G12 = fma2(NB, R8, R5);
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
        const float a00 = lo(A00) + hi(A00), a01 = lo(A01) + hi(A01), a02 = lo(A02) + hi(A02);
        const float a11 = lo(A11) + hi(A11), a12 = lo(A12) + hi(A12), a22 = lo(A22) + hi(A22);
        const float b0 = lo(B0) + hi(B0), b1 = lo(B1) + hi(B1), b2 = lo(B2) + hi(B2);
        const float cst = lo(CST) + hi(CST);
        S.costp[f][l] = cst;
        if (PERSIST) {
            // Il aliases the staged outputs of the previous tile: its bulk stores must have read them; then the
            // input buffer is free and is refilled for the next tile while P2b..P5 of this one run
            if (tid == 0 && out_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            if (bulk_in && tid == 0) {
                const int tn = tile + gridDim.x;
                if (tn < n_tiles && (tn + 1) * FT <= n_frames) issue_mw(tn);
            }
        } else if (staged) {
            __syncthreads();           // the input tiles alias Il: every thread is done reading them
        }
        PHASE_MARK(3);
        // spatial inertia of this marker about the head point: B^T A B with B = [-[p]x  I]
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
        float4* dst = reinterpret_cast<float4*>(S.Il[tid]);
#pragma unroll
        for (int i = (WANT_H ? 0 : 5); i < (NSP + 1) / 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    __syncthreads();
    PHASE_MARK(4);

    // ---- P3: subtree sums up the kinematic tree, one thread per (component pair, frame), packed adds
    {
        constexpr int NKP = (NSP + 1) / 2;       // 14 component pairs
        const int task = tid;
        const int f = task / NKP;
        const int kp = task - f * NKP;
        if (task < FT * NKP && (WANT_H || kp >= 10)) {
            f2 v[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) v[l] = pk(*reinterpret_cast<const float2*>(&S.Il[f * NL + l][2 * kp]));
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
    __syncthreads();   // Il is dead from here on; the staged outputs alias it
    PHASE_MARK(5);

    // ---- P4a: y_beta = I_subtree(beta) tau_beta, g, translation rows.  task (beta, frame);
    //      beta == 22 is the translation block
    for (int task = tid; task < FT * (NANG + 1); task += NT) {
        const int be = task / FT;
        const int f = task - be * FT;
        float* g = S.o.g[f];
        float* H = S.o.H[f];
        if (be == NANG) {
            float c = 0.f;
#pragma unroll
            for (int l = 0; l < NL; ++l) c += S.costp[f][l];
            S.o.cost[f] = c;
            const float* I0 = S.Ij[f];
            g[0] = I0[24]; g[1] = I0[25]; g[2] = I0[26];
            if (WANT_H) {
                H[upper_index(0, 0)] = I0[15]; H[upper_index(0, 1)] = I0[16]; H[upper_index(0, 2)] = I0[17];
                H[upper_index(1, 1)] = I0[18]; H[upper_index(1, 2)] = I0[19]; H[upper_index(2, 2)] = I0[20];
            }
            continue;
        }
        const float4* I4 = reinterpret_cast<const float4*>(&S.Ij[f][c_tab.joint[be] * (NSP + 1)]);
        const float4 i0 = I4[0], i1 = I4[1], i2 = I4[2], i3 = I4[3], i4 = I4[4], i5 = I4[5], i6 = I4[6];
        const float I[28] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w, i3.x, i3.y,
                             i3.z, i3.w, i4.x, i4.y, i4.z, i4.w, i5.x, i5.y, i5.z, i5.w, i6.x, i6.y, i6.z, i6.w};
        const float4 t0 = *reinterpret_cast<const float4*>(&S.tau[f][be * TAU_STRIDE]);
        const float2 t1 = *reinterpret_cast<const float2*>(&S.tau[f][be * TAU_STRIDE + 4]);
        const float o0 = t0.x, o1 = t0.y, o2 = t0.z, v0 = t0.w, v1 = t1.x, v2 = t1.y;
        g[3 + be] = o0 * I[21] + o1 * I[22] + o2 * I[23] + v0 * I[24] + v1 * I[25] + v2 * I[26];
        if (!WANT_H) continue;
        // y = I tau_beta ; I = [[TL, PA],[PA^T, A]]
        const float yt0 = I[0] * o0 + I[1] * o1 + I[2] * o2 + I[6] * v0 + I[7] * v1 + I[8] * v2;
        const float yt1 = I[1] * o0 + I[3] * o1 + I[4] * o2 + I[9] * v0 + I[10] * v1 + I[11] * v2;
        const float yt2 = I[2] * o0 + I[4] * o1 + I[5] * o2 + I[12] * v0 + I[13] * v1 + I[14] * v2;
        const float yb0 = I[6] * o0 + I[9] * o1 + I[12] * o2 + I[15] * v0 + I[16] * v1 + I[17] * v2;
        const float yb1 = I[7] * o0 + I[10] * o1 + I[13] * o2 + I[16] * v0 + I[18] * v1 + I[19] * v2;
        const float yb2 = I[8] * o0 + I[11] * o1 + I[14] * o2 + I[17] * v0 + I[19] * v1 + I[20] * v2;
        *reinterpret_cast<float4*>(&S.o.y[f][be * TAU_STRIDE]) = make_float4(yt0, yt1, yt2, yb0);
        *reinterpret_cast<float2*>(&S.o.y[f][be * TAU_STRIDE + 4]) = make_float2(yb1, yb2);
        const int sb = 3 + be;  // active slot of beta: rows 0..2 (translation) of column sb
        H[sb] = yb0;
        H[NA + sb - 1] = yb1;
        H[2 * NA + sb - 3] = yb2;
    }
    __syncthreads();
    PHASE_MARK(6);

    // ---- P4b: one entry per (angle pair, frame): H[al][be] = tau_al . y_be.  NT / FT = 20 exactly, so a
    //      thread keeps its frame (tid % FT) and walks the table with stride 20: p = tid / FT + 20 k.
    //      All loads and dot products first, all stores last: no store -> load ordering between entries.
    if (WANT_H) {
        constexpr int NE = (N_REL + NL - 1) / NL;    // 10 related entries per thread
        const int f = tid % FT, p0 = tid / FT;
        const float* tau_f = &S.tau[f][0];
        const float* y_f = &S.o.y[f][0];
        float* H_f = &S.o.H[f][0];
        unsigned e[NE];
        float hv[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) e[k] = S.tab[min(p0 + NL * k, N_REL - 1)];
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(tau_f + (e[k] & 0xFFu));
            const float2 a1 = *reinterpret_cast<const float2*>(tau_f + (e[k] & 0xFFu) + 4);
            const float4 y0 = *reinterpret_cast<const float4*>(y_f + ((e[k] >> 8) & 0xFFu));
            const float2 y1 = *reinterpret_cast<const float2*>(y_f + ((e[k] >> 8) & 0xFFu) + 4);
            hv[k] = a0.x * y0.x + a0.y * y0.y + a0.z * y0.z + a0.w * y0.w + a1.x * y1.x + a1.y * y1.y;
        }
#pragma unroll
        for (int k = 0; k < NE; ++k)
            if (p0 + NL * k < N_REL) H_f[e[k] >> 16] = hv[k];
        // structural zeros: (N_PAIR - N_REL) x FT stores
#pragma unroll
        for (int k = 0; k < ((N_PAIR - N_REL) * FT + NT - 1) / NT; ++k) {
            const int z = p0 + NL * k;
            if (z < N_PAIR - N_REL) H_f[S.tab[N_REL + z] >> 16] = 0.f;
        }
    }
    __syncthreads();
    PHASE_MARK(7);

    // ---- P5: write-out.  The staged blocks have the global layout.  Full tiles with 16-byte aligned outputs:
    //      three bulk async stores (TMA) issued by one thread; otherwise straight vector copies
    if ((use_bulk & 2) && nf == FT) {
        if (tid == 0) {
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
    }   // tiles
    if (tid == 0 && out_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// register budget by "at least MINB resident CTAs" ...
template <int FT, bool WANT_H, int MINB, int NPAIR, bool PERSIST>
__global__ void __launch_bounds__(FT * NL, MINB)
fte_eval_kernel(const __grid_constant__ SceneF scene, const int n_frames, const int use_bulk,
                const float* __restrict__ xg, const float* __restrict__ meas, const float* __restrict__ wts,
                float* __restrict__ cost_out, float* __restrict__ g_out, float* __restrict__ H_out) {
    fte_eval_body<FT, WANT_H, NPAIR, PERSIST>(scene, n_frames, use_bulk, xg, meas, wts, cost_out, g_out, H_out);
}
#ifdef ACINO_EXPERIMENTS
// ... or by an explicit register cap (5 CTAs x 160 threads x 80 registers = 64 000 of the SM's 65 536)
template <int FT, bool WANT_H, int MAXREG, int NPAIR, bool PERSIST>
__global__ void __launch_bounds__(FT * NL) __maxnreg__(MAXREG)
fte_eval_kernel_r(const __grid_constant__ SceneF scene, const int n_frames, const int use_bulk,
                  const float* __restrict__ xg, const float* __restrict__ meas, const float* __restrict__ wts,
                  float* __restrict__ cost_out, float* __restrict__ g_out, float* __restrict__ H_out) {
    fte_eval_body<FT, WANT_H, NPAIR, PERSIST>(scene, n_frames, use_bulk, xg, meas, wts, cost_out, g_out, H_out);
}
#endif

#ifdef ACINO_PHASE_TIMING
extern "C" void acino_debug_phase_cycles(long long* out16) { cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(long long) * 16); }
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
using FteKernel = void (*)(const SceneF, int, int, const float*, const float*, const float*, float*, float*, float*);

// per-device launch configuration (cudaFuncSetAttribute is per device; a process may drive several GPUs)
struct FteDeviceCfg {
    FteKernel configured[16];
    int n_configured;
    int n_sm;
};
static FteDeviceCfg g_dev_cfg[64];

template <int FT, bool PERSIST>
static cudaError_t launch_fte_eval_k(FteKernel kH, FteKernel kN, int ctas_per_sm, const SceneF& scene, int n_frames,
                                     const float* x, const float* meas, const float* w, float* cost, float* g, float* H,
                                     cudaStream_t stream, int exp_bits, size_t smem_pad) {
    // bulk (TMA) staging needs 16-byte aligned tiles: frame tiles are multiples of 16 bytes, so it is the
    // base pointers that decide
    // bit 0: inputs, bit 1: outputs (cost tiles are FT * 4 = 32 bytes, g / H tiles multiples of 16 bytes)
    int use_bulk = (((((uintptr_t)x | (uintptr_t)meas | (uintptr_t)w) & 15u) == 0) ? 1 : 0) |
                   (((((uintptr_t)cost | (uintptr_t)g | (uintptr_t)H) & 15u) == 0 && FT % 4 == 0) ? 2 : 0);
    if ((exp_bits & 1) && (use_bulk & 1)) use_bulk |= 4;      // experiment: L2 prefetch of a later tile
    if (exp_bits & 2) use_bulk |= 8;                          // experiment: MUFU sin / cos
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    FteDeviceCfg& cfg = g_dev_cfg[dev];
    const int n_tiles = (n_frames + FT - 1) / FT;
    int grid = n_tiles;
    if (PERSIST) {       // one wave of resident CTAs, each walking tiles blockIdx.x, blockIdx.x + grid, ...
        if (!cfg.n_sm) cudaDeviceGetAttribute(&cfg.n_sm, cudaDevAttrMultiProcessorCount, dev);
        grid = n_tiles < cfg.n_sm * ctas_per_sm ? n_tiles : cfg.n_sm * ctas_per_sm;
    }
    const size_t smem = sizeof(Smem<FT, PERSIST>) + smem_pad;
    FteKernel k = H ? kH : kN;
    bool done = false;
    for (int i = 0; i < cfg.n_configured; ++i) done |= cfg.configured[i] == k;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cfg.n_configured < 16) cfg.configured[cfg.n_configured++] = k;
    }
    k<<<grid, FT * NL, smem, stream>>>(scene, n_frames, use_bulk, x, meas, w, cost, g, H);
    return cudaGetLastError();
}

template <int FT, int MINB, int NPAIR, bool PERSIST>
static cudaError_t launch_fte_eval_v(const SceneF& scene, int n_frames, const float* x, const float* meas,
                                     const float* w, float* cost, float* g, float* H, cudaStream_t stream, int exp_bits = 0,
                                     size_t smem_pad = 0) {
    return launch_fte_eval_k<FT, PERSIST>(fte_eval_kernel<FT, true, MINB, NPAIR, PERSIST>, fte_eval_kernel<FT, false, MINB, NPAIR, PERSIST>,
                                          MINB, scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, smem_pad);
}
#ifdef ACINO_EXPERIMENTS
template <int FT, int MAXREG, int CTAS, int NPAIR, bool PERSIST>
static cudaError_t launch_fte_eval_r(const SceneF& scene, int n_frames, const float* x, const float* meas,
                                     const float* w, float* cost, float* g, float* H, cudaStream_t stream, int exp_bits,
                                     size_t smem_pad) {
    return launch_fte_eval_k<FT, PERSIST>(fte_eval_kernel_r<FT, true, MAXREG, NPAIR, PERSIST>, fte_eval_kernel_r<FT, false, MAXREG, NPAIR, PERSIST>,
                                          CTAS, scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, smem_pad);
}
#endif

cudaError_t launch_fte_eval(const SceneF& scene, int n_frames, const float* x, const float* meas,
                            const float* w, float* cost, float* g, float* H, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
#ifdef ACINO_EXPERIMENTS
    // A/B variants of scripts/bench_variants.sh (B200, 256 000 frames, profiles/r01_fte_eval.md); build with
    // -DACINO_EXPERIMENTS to get the ACINO_FTE_VARIANT / ACINO_FTE_EXP / ACINO_FTE_SMEM_PAD switches:
    //   0 default 6.45e8 frames/s | 1: 5 CTAs/SM, 72 regs 6.2e8 | 4: 16 frames/CTA 4.8e8 | 5/6: 4 frames/CTA 5.7e8 / 5.4e8
    //   8: camera-pair loop unrolled 6.4e8 | 9: unrolled, 3 CTAs/SM, 128 regs 5.6e8 | 10: persistent 5.9e8
    static int variant = -1, exp_bits = 0;
    static size_t pad = 0;
    if (variant < 0) {
        const char* e = getenv("ACINO_FTE_VARIANT");
        variant = e ? atoi(e) : 0;
        e = getenv("ACINO_FTE_EXP");
        exp_bits = e ? atoi(e) : 0;
        e = getenv("ACINO_FTE_SMEM_PAD");
        pad = e ? (size_t)atol(e) : 0;
    }
    if (variant == 1) return launch_fte_eval_v<8, 5, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 4) return launch_fte_eval_v<16, 2, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 5) return launch_fte_eval_v<4, 9, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 6) return launch_fte_eval_v<4, 10, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 8 && scene.n_cams == 6) return launch_fte_eval_v<8, 4, 3, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 9 && scene.n_cams == 6) return launch_fte_eval_v<8, 3, 3, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 10) return launch_fte_eval_v<8, 4, 0, true>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 11) return launch_fte_eval_r<8, 80, 5, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    if (variant == 12) return launch_fte_eval_r<8, 88, 4, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
    return launch_fte_eval_v<8, 4, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream, exp_bits, pad);
#else
    // 8 frames per CTA, 4 CTAs per SM (96 registers, no spills), runtime camera-pair loop
    return launch_fte_eval_v<8, 4, 0, false>(scene, n_frames, x, meas, w, cost, g, H, stream);
#endif
}

const char* fte_eval_kernel_name(int) { return "fte_eval_kernel<8, 1, 4, 0, 0>"; }

cudaError_t launch_fk_project(const SceneF& scene, int n_frames, const float* x, float* pos, float* uv,
                              cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int grid = (n_frames + FTP - 1) / FTP;
    fk_project_kernel<<<grid, FTP * NL, 0, stream>>>(scene, n_frames, x, pos, uv);
    return cudaGetLastError();
}

}  // namespace acino
