// Shared device-side definitions for libacino_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/acino_b200.h"

namespace acino {

constexpr int NA = ACINO_N_ACTIVE;    // 25 active pose parameters
constexpr int NL = ACINO_N_MARKERS;   // 20 markers
constexpr int NU = ACINO_N_UPPER;     // 325
constexpr int NANG = 22;              // angle slots (active slots 3..24)
constexpr int NJ = 14;                // joints in the rotation chain
constexpr int NSP = 27;               // 21 spatial-inertia + 6 wrench components

// One camera, fp32: world->camera rotation (row-major), translation, intrinsics, distortion.
// Kannala-Brandt model of pt3d_to_2d (reference all_optimizations.py:193-209).
struct CamF {
    float R[9];
    float t[3];
    float fx, fy, cx, cy;
    float D[4];
    float D3[4];          // 3 k1, 5 k2, 7 k3, 9 k4 (derivative of the distortion polynomial)
};

struct CamD {
    double R[9];
    double t[3];
    double fx, fy, cx, cy;
    double D[4];
};

struct LossF {
    float a, b, c;        // break points
    float ea, eb, ec;     // exp(a), exp(b), exp(c)
    float p2c;            // -a^2/2
    float k3;             // a (c-b) / 2
    float inv_cb;         // 1/(c-b)
    float p3c;            // a b - a^2/2
    float p4;             // a b - a^2/2 + a (c-b)/2
    float rho0;           // rho(0)
    float kappa;          // a / (2 (c - b))
    float m2kappa;        // -2 kappa
};

// Two cameras (2k, 2k+1) interleaved as float2 pairs: operands of the packed fp32 (FFMA2) path
struct CamPairF {
    float2 R[9];
    float2 t[3];
    float2 fx, fy, cx, cy;
    float2 D[4];
    float2 D3[4];
};

struct SceneF {
    CamF cam[ACINO_MAX_CAMS];
    CamPairF pair[ACINO_MAX_CAMS / 2];
    LossF loss;
    int n_cams;
};

__host__ __device__ inline int upper_index(int i, int j) {  // i <= j
    return i * NA - (i * (i - 1)) / 2 + (j - i);
}

// ------------------------------------------------------------------------------------------
// Fisheye projection of a camera-frame point with its 2x3 Jacobian w.r.t. the camera-frame
// point (closed form).  T = float or double.
template <typename T>
struct ProjOut {
    T u, v;        // pixel WITHOUT the principal point: u = fx a s, v = fy b s
    T ju[3];       // d u / d Xc
    T jv[3];       // d v / d Xc
};

template <typename T> __device__ __forceinline__ T t_rsqrt(T x);
template <> __device__ __forceinline__ float t_rsqrt<float>(float x) { return rsqrtf(x); }
template <> __device__ __forceinline__ double t_rsqrt<double>(double x) { return 1.0 / sqrt(x); }
template <typename T> __device__ __forceinline__ T t_atan(T x);
template <> __device__ __forceinline__ float t_atan<float>(float x) { return atanf(x); }
template <> __device__ __forceinline__ double t_atan<double>(double x) { return atan(x); }
template <typename T> __device__ __forceinline__ T t_rcp(T x);
__device__ __forceinline__ float fast_rcp(float x) {  // MUFU.RCP, ~1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <> __device__ __forceinline__ float t_rcp<float>(float x) { return fast_rcp(x); }
template <> __device__ __forceinline__ double t_rcp<double>(double x) { return 1.0 / x; }

template <typename T, bool WITH_JAC>
__device__ __forceinline__ void fisheye_cam(const T x, const T y, const T z, const T fx, const T fy,
                                            const T k1, const T k2, const T k3, const T k4,
                                            ProjOut<T>& o) {
    const T iz = t_rcp<T>(z);
    const T a = x * iz;
    const T b = y * iz;
    const T r2 = a * a + b * b + T(1e-12);     // the "+1e-12" of all_optimizations.py:201
    const T ir = t_rsqrt<T>(r2);
    const T r = r2 * ir;
    const T th = t_atan<T>(r);
    const T th2 = th * th;
    const T td = th * (T(1) + th2 * (k1 + th2 * (k2 + th2 * (k3 + th2 * k4))));
    const T s = td * ir;
    o.u = fx * (a * s);
    o.v = fy * (b * s);
    if (WITH_JAC) {
        const T dtd = T(1) + th2 * (T(3) * k1 + th2 * (T(5) * k2 + th2 * (T(7) * k3 + th2 * (T(9) * k4))));
        // q = (ds/dr)/r = (dtd/(1+r^2) - s) / r^2
        const T q = (dtd * t_rcp<T>(T(1) + r2) - s) * (ir * ir);
        const T aq = a * q;
        const T m00 = s + a * aq;
        const T m01 = b * aq;
        const T m11 = s + b * b * q;
        const T fxi = fx * iz;
        const T fyi = fy * iz;
        o.ju[0] = fxi * m00;
        o.ju[1] = fxi * m01;
        o.ju[2] = -fxi * (m00 * a + m01 * b);
        o.jv[0] = fyi * m01;
        o.jv[1] = fyi * m11;
        o.jv[2] = -fyi * (m01 * a + m11 * b);
    }
}

// ------------------------------------------------------------------------------------------
// Redescending loss (reference build.py:382-395, literal logistic blend) for e = |w r| >= 0:
// returns rho(e), rho'(e) and the Gauss-Newton curvature weight
//     psi(e) = max(rho'(e)/e, 1 - sigma_a(e))
// i.e. the IRLS weight rho'/e floored by the (frozen-gate) curvature of the quadratic piece.
// The floor takes over for e < ~0.45, where the literal blend has a tiny cusp (rho'(0+) < 0)
// that makes rho'/e both negative and ill-conditioned in fp32; psi >= 0 everywhere.
__device__ __forceinline__ void redescending(const LossF& L, const float e, float& rho, float& drho,
                                             float& psi) {
    const float E = __expf(-e);
    const float sa = fast_rcp(fmaf(E, L.ea, 1.0f));
    const float sb = fast_rcp(fmaf(E, L.eb, 1.0f));
    const float sc = fast_rcp(fmaf(E, L.ec, 1.0f));
    const float dsa = sa * (1.0f - sa);
    const float dsb = sb * (1.0f - sb);
    const float dsc = sc * (1.0f - sc);
    const float p1 = 0.5f * e * e;
    const float p2 = fmaf(L.a, e, L.p2c);
    const float u = (L.c - e) * L.inv_cb;
    const float p3 = fmaf(L.k3, 1.0f - u * u, L.p3c);
    const float dp3 = L.a * u;
    const float gab = sa - sb;
    const float gbc = sb - sc;
    rho = (1.0f - sa) * p1 + gab * p2 + gbc * p3 + sc * L.p4;
    drho = (1.0f - sa) * e - dsa * p1 + (dsa - dsb) * p2 + gab * L.a + (dsb - dsc) * p3 + gbc * dp3 +
           dsc * L.p4;
    psi = e > 0.0f ? fmaxf(__fdividef(drho, e), 1.0f - sa) : 1.0f - sa;
}

// ------------------------------------------------------------------------------------------
// Fused-kernel variants (fp32, instruction-count optimised; same arithmetic, fewer operations).

__device__ __forceinline__ float fast_rsqrt(float x) {   // MUFU.RSQ, no denormal fix-up (x >= 1e-12 here)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_ex2(float x) {     // MUFU.EX2 without exp2f's denormal-range rescaling
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// atan(r) for r >= 0 given ir = 1/r: degree-7 polynomial in z^2 on z = min(r, 1/r) in [0,1]
// (Chebyshev-node fit, max abs error 1.2e-7), atan(r) = pi/2 - atan(1/r) for r > 1.
__device__ __forceinline__ float atan_pos(float r, float ir) {
    const float z = fminf(r, ir);
    const float s = z * z;
    float p = 0.003866738872602582f;
    p = fmaf(p, s, -0.02002674713730812f);
    p = fmaf(p, s, 0.04891432076692581f);
    p = fmaf(p, s, -0.08009681850671768f);
    p = fmaf(p, s, 0.1086575910449028f);
    p = fmaf(p, s, -0.14257045090198517f);
    p = fmaf(p, s, 0.19998681545257568f);
    p = fmaf(p, s, -0.33333322405815125f);
    p = p * s;
    p = fmaf(p, z, z);
    return r > 1.0f ? 1.5707963267948966f - p : p;
}

// Redescending loss, algebraically regrouped (exactly the literal blend of build.py:388-395):
//   rho  = e^2/2 - sa ta^2/2 - kappa sb tb^2 + kappa sc tc^2,   t* = e - {a,b,c}, kappa = a/(2(c-b))
//   rho' = e - sa ta - 2 kappa (sb tb - sc tc) - S1 + sa(sa ta^2)/2 ... (product rule on the gates)
// Returns rho, psi_raw = rho'(e)/e (unclamped, used for the gradient: d rho/d r = psi_raw w^2 r) and
// the curvature floor 1 - sigma_a.  e must be in [tiny, 40].
__device__ __forceinline__ void redescending_fast(const LossF& L, const float e, float& rho, float& psi_raw,
                                                  float& floor_) {
    const float E = __expf(-e);
    const float sa = fast_rcp(fmaf(E, L.ea, 1.0f));
    const float sb = fast_rcp(fmaf(E, L.eb, 1.0f));
    const float sc = fast_rcp(fmaf(E, L.ec, 1.0f));
    const float ta = e - L.a, tb = e - L.b, tc = e - L.c;
    const float qa = sa * ta, qb = sb * tb, qc = sc * tc;
    const float ra = qa * ta, rb = qb * tb, rc = qc * tc;
    const float S1 = fmaf(L.kappa, rb - rc, 0.5f * ra);
    rho = fmaf(0.5f * e, e, -S1);
    float d = e - qa;
    d = fmaf(L.m2kappa, qb - qc, d);
    d -= S1;
    float w1 = sb * rb;
    w1 = fmaf(-sc, rc, w1);
    d = fmaf(L.kappa, w1, d);
    d = fmaf(0.5f * sa, ra, d);
    psi_raw = d * fast_rcp(e);
    floor_ = 1.0f - sa;
}

// ------------------------------------------------------------------------------------------
// Packed fp32 pairs: Blackwell (sm_100) issues fma/add/mul on two fp32 values per instruction
// (PTX fma.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2).  Same FMA-pipe throughput as two scalar FFMAs but
// half the issue slots - and this kernel is issue-bound.  A broadcast pair (s, s) is encoded by ptxas
// as a scalar operand, so per-thread scalars and constants cost nothing extra.
struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 pk(float2 a) { return pk(a.x, a.y); }
__device__ __forceinline__ f2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(f2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float hi(f2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ f2 rcp2(f2 a) { return pk(fast_rcp(lo(a)), fast_rcp(hi(a))); }
__device__ __forceinline__ f2 rsqrt2(f2 a) { return pk(fast_rsqrt(lo(a)), fast_rsqrt(hi(a))); }

// atan of a pair r >= 0 given ir = 1/r (see atan_pos)
__device__ __forceinline__ f2 atan_pos2(f2 r, f2 ir) {
    const f2 z = pk(fminf(lo(r), lo(ir)), fminf(hi(r), hi(ir)));
    const f2 s = mul2(z, z);
    f2 p = bc(0.003866738872602582f);
    p = fma2(p, s, bc(-0.02002674713730812f));
    p = fma2(p, s, bc(0.04891432076692581f));
    p = fma2(p, s, bc(-0.08009681850671768f));
    p = fma2(p, s, bc(0.1086575910449028f));
    p = fma2(p, s, bc(-0.14257045090198517f));
    p = fma2(p, s, bc(0.19998681545257568f));
    p = fma2(p, s, bc(-0.33333322405815125f));
    p = mul2(p, s);
    p = fma2(p, z, z);
    const float pl = lo(p), ph = hi(p);
    return pk(lo(r) > 1.0f ? 1.5707963267948966f - pl : pl, hi(r) > 1.0f ? 1.5707963267948966f - ph : ph);
}

// redescending_fast on a pair of e values in [tiny, 40]
__device__ __forceinline__ void redescending_fast2(const LossF& L, const f2 e, f2& rho, f2& psi_raw, f2& floor_) {
    const f2 ne = mul2(e, bc(-1.4426950408889634f));
    const f2 E = pk(fast_ex2(lo(ne)), fast_ex2(hi(ne)));       // e <= 40: 2^-58, far from the denormal range exp2f guards
    const f2 one = bc(1.0f);
    const f2 sa = rcp2(fma2(E, bc(L.ea), one));
    const f2 sb = rcp2(fma2(E, bc(L.eb), one));
    const f2 sc = rcp2(fma2(E, bc(L.ec), one));
    const f2 ta = sub2(e, bc(L.a)), tb = sub2(e, bc(L.b)), tc = sub2(e, bc(L.c));
    const f2 qa = mul2(sa, ta), qb = mul2(sb, tb), qc = mul2(sc, tc);
    const f2 ra = mul2(qa, ta), rb = mul2(qb, tb), rc = mul2(qc, tc);
    const f2 S1 = fma2(bc(L.kappa), sub2(rb, rc), mul2(ra, bc(0.5f)));
    rho = sub2(mul2(mul2(e, bc(0.5f)), e), S1);
    f2 d = sub2(e, qa);
    d = fma2(bc(L.m2kappa), sub2(qb, qc), d);
    d = sub2(d, S1);
    const f2 w1 = sub2(mul2(sb, rb), mul2(sc, rc));
    d = fma2(bc(L.kappa), w1, d);
    d = fma2(mul2(sa, bc(0.5f)), ra, d);
    psi_raw = mul2(d, rcp2(e));
    floor_ = sub2(one, sa);
}

}  // namespace acino
