// Cheetah forward kinematics shared by the fused FTE kernels (fte_eval.cu, fte_jac.cu): rotation chain
// RI_0..RI_13 (reference all_optimizations.py:101-128), marker offsets (:138-165) and the world-frame
// twist (omega, pivot x omega) of every angle, d p_l / d angle = omega x p_l + v for every marker l
// carried by a joint at or below the angle's joint.
#pragma once
#include "acino_common.cuh"

namespace acino {

constexpr int TAU_STRIDE = 8;           // (omega, v) padded to 8 floats: one LDS.128 + one LDS.64
constexpr int TAUF = (NANG + 1) * TAU_STRIDE + 4;  // 22 twists + one all-zero slot; stride 188 = 28 (mod 32): conflict-free across frames

// joint of each angle slot (angle slot s <-> active slot 3+s):
// phi0 phi1 phi3 | theta0..13 | psi0 psi1 psi3 psi4 psi5
constexpr int k_angle_joint[NANG] = {0, 1, 3, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0, 1, 3, 4, 5};
// parent of each joint in the rotation chain (all_optimizations.py:101-128)
constexpr int k_joint_parent[NJ] = {-1, 0, 1, 2, 3, 4, 2, 6, 2, 8, 3, 10, 3, 12};

constexpr bool joint_is_anc(int a, int k) {  // a ancestor-or-self of k
    while (k >= 0) {
        if (k == a) return true;
        k = k_joint_parent[k];
    }
    return false;
}

// joint that carries each marker (order of all_optimizations.py:170-178)
constexpr int k_marker_joint[NL] = {0, 0, 0, 1, 2, 3, 4, 5, 2, 6, 7, 2, 8, 9, 3, 10, 11, 3, 12, 13};

struct Col3 {
    float x, y, z;
};
__device__ __forceinline__ Col3 operator*(float s, Col3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ Col3 operator+(Col3 a, Col3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Col3 operator-(Col3 a, Col3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Col3 fma3(float s, Col3 a, Col3 b) {
    return {fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)};
}
__device__ __forceinline__ float fma3(float s, float a, float b) { return fmaf(s, a, b); }
__device__ __forceinline__ Col3 cross(Col3 a, Col3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Body->world rotation R_k_I stored by columns.  V = Col3: the whole matrix in one thread.  V = float: ONE ROW of it
// (c0, c1, c2 are then the three entries of row i) - right-multiplication by an elementary rotation mixes two columns,
// so the three rows of the chain are independent and three threads can each carry one row of the same frame.
template <typename V>
struct Mat3T {
    V c0, c1, c2;
};
using Mat3 = Mat3T<Col3>;
// M <- M * Ry_a(th):  (c0,c2) <- (c c0 - s c2, s c0 + c c2)       [rot_y transposed, :75-82]
template <typename V>
__device__ __forceinline__ void rot_y(Mat3T<V>& M, float s, float c) {
    const V a = M.c0, b = M.c2;
    M.c0 = fma3(c, a, (-s) * b);
    M.c2 = fma3(s, a, c * b);
}
// M <- M * Rx_a(ph):  (c1,c2) <- (c c1 + s c2, -s c1 + c c2)      [rot_x transposed, :66-73]
template <typename V>
__device__ __forceinline__ void rot_x(Mat3T<V>& M, float s, float c) {
    const V a = M.c1, b = M.c2;
    M.c1 = fma3(c, a, s * b);
    M.c2 = fma3(-s, a, c * b);
}
// M <- M * Rz_a(ps):  (c0,c1) <- (c c0 + s c1, -s c0 + c c1)      [rot_z transposed, :84-91]
template <typename V>
__device__ __forceinline__ void rot_z(Mat3T<V>& M, float s, float c) {
    const V a = M.c0, b = M.c1;
    M.c0 = fma3(c, a, s * b);
    M.c1 = fma3(-s, a, c * b);
}

struct FkWriter {
    float* p;    // [NL][3]
    float* tau;  // [NANG][TAU_STRIDE]
    __device__ __forceinline__ void marker(int l, Col3 v) const {
        p[l * 3 + 0] = v.x;
        p[l * 3 + 1] = v.y;
        p[l * 3 + 2] = v.z;
    }
    // twist of angle slot s: axis omega (world), pivot (relative to head)
    __device__ __forceinline__ void twist(int s, Col3 om, Col3 piv) const {
        const Col3 v = cross(piv, om);
        float* t = tau + s * TAU_STRIDE;
        t[0] = om.x; t[1] = om.y; t[2] = om.z;
        t[3] = v.x;  t[4] = v.y;  t[5] = v.z;
    }
    __device__ __forceinline__ void twist0(int s, Col3 om) const {  // pivot = head
        float* t = tau + s * TAU_STRIDE;
        t[0] = om.x; t[1] = om.y; t[2] = om.z;
        t[3] = 0.f;  t[4] = 0.f;  t[5] = 0.f;
    }
};

// Row writer (V = float): thread i of a frame stores component i of every marker and of every rotation axis.  The
// linear part v = pivot x omega of a twist needs all three components, so it is formed later from the stored values
// (every pivot is a marker position: k_angle_pivot).
struct FkRowWriter {
    float* p;    // [NL][3] + i
    float* tau;  // [NANG][TAU_STRIDE] + i
    __device__ __forceinline__ void marker(int l, float v) const { p[l * 3] = v; }
    __device__ __forceinline__ void twist(int s, float om, float) const { tau[s * TAU_STRIDE] = om; }
    __device__ __forceinline__ void twist0(int s, float om) const { tau[s * TAU_STRIDE] = om; }
};

// angle slot ids
enum { A_PHI0 = 0, A_PHI1 = 1, A_PHI3 = 2, A_TH0 = 3, A_PSI0 = 17, A_PSI1 = 18, A_PSI3 = 19, A_PSI4 = 20, A_PSI5 = 21 };

// marker whose position is the pivot of each angle slot's rotation (-1: the head point, v = 0); matches the
// w.twist(...) calls of cheetah_fk_t below
constexpr int k_angle_pivot[NANG] = {-1, -1, 4,                                      // phi0 phi1 phi3
                                     -1, -1, 3, 4, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18,  // theta0..13
                                     -1, -1, 4, 5, 6};                               // psi0 psi1 psi3 psi4 psi5

// Cheetah forward kinematics of one frame: marker positions relative to the head point and the
// world-frame twist of every angle.  Follows the chain RI_0..RI_13 (:101-128) and p_* (:138-165);
// R_k_I = R_parent_I Ry_a(theta) Rx_a(phi) Rz_a(psi), and the world axis of each angle is the
// matching column of the partially composed matrix (theta: column 1 before Ry; phi: column 0
// after Ry; psi: column 2 after Rx).  `M` enters as the identity (V = Col3) or as row i of it (V = float).
// SEG selects which outputs are written (the arithmetic nothing written depends on is eliminated by the compiler):
// bit 0 head / neck / torso / tail (markers 0..7 and their angles), bit 1 front legs (markers 8..13, theta6..9), bit 2 back
// legs (markers 14..19, theta10..13) - three warps can each run one segment of the same frames.
template <typename V, int SEG, typename W>
__device__ __forceinline__ void cheetah_fk_t(const float2* __restrict__ sc, Mat3T<V> M, const W& w) {
    constexpr bool S0 = SEG & 1, S1 = SEG & 2, S2 = SEG & 4;
    float sn[NANG], cs[NANG];
#pragma unroll
    for (int i = 0; i < NANG; ++i) {
        const float2 v = sc[i];
        sn[i] = v.x;
        cs[i] = v.y;
    }
#define TH(k) sn[A_TH0 + (k)], cs[A_TH0 + (k)]
    // joint 0: head
    if (S0) w.twist0(A_TH0 + 0, M.c1);
    rot_y(M, TH(0));
    if (S0) w.twist0(A_PHI0, M.c0);
    rot_x(M, sn[A_PHI0], cs[A_PHI0]);
    if (S0) w.twist0(A_PSI0, M.c2);
    rot_z(M, sn[A_PSI0], cs[A_PSI0]);
    if (S0) w.marker(0, 0.03f * M.c1);                          // l_eye
    if (S0) w.marker(1, -0.03f * M.c1);                         // r_eye
    if (S0) w.marker(2, 0.055f * (M.c0 - M.c2));                // nose
    // joint 1: neck
    if (S0) w.twist0(A_TH0 + 1, M.c1);
    rot_y(M, TH(1));
    if (S0) w.twist0(A_PHI1, M.c0);
    rot_x(M, sn[A_PHI1], cs[A_PHI1]);
    if (S0) w.twist0(A_PSI1, M.c2);
    rot_z(M, sn[A_PSI1], cs[A_PSI1]);
    const V neck = -0.28f * M.c0;
    if (S0) w.marker(3, neck);
    // joint 2: front torso (pivot neck_base)
    if (S0) w.twist(A_TH0 + 2, M.c1, neck);
    rot_y(M, TH(2));
    const Mat3T<V> M2 = M;
    const V spine = fma3(-0.37f, M2.c0, neck);
    if (S0) w.marker(4, spine);
    const V sh_mid = fma3(-0.04f, M2.c0, fma3(-0.10f, M2.c2, neck));
    const V lsh = fma3(0.08f, M2.c1, sh_mid);
    const V rsh = fma3(-0.08f, M2.c1, sh_mid);
    if (S1) w.marker(8, lsh);
    if (S1) w.marker(11, rsh);
    // joints 6,7: left front leg; 8,9: right front leg (axis = column 1 of M2 throughout)
    {
        Mat3T<V> Ml = M2;
        if (S1) w.twist(A_TH0 + 6, M2.c1, lsh);
        rot_y(Ml, TH(6));
        const V knee = fma3(-0.24f, Ml.c2, lsh);
        if (S1) w.marker(9, knee);
        if (S1) w.twist(A_TH0 + 7, M2.c1, knee);
        rot_y(Ml, TH(7));
        if (S1) w.marker(10, fma3(-0.28f, Ml.c2, knee));
        Mat3T<V> Mr = M2;
        if (S1) w.twist(A_TH0 + 8, M2.c1, rsh);
        rot_y(Mr, TH(8));
        const V kneer = fma3(-0.24f, Mr.c2, rsh);
        if (S1) w.marker(12, kneer);
        if (S1) w.twist(A_TH0 + 9, M2.c1, kneer);
        rot_y(Mr, TH(9));
        if (S1) w.marker(13, fma3(-0.28f, Mr.c2, kneer));
    }
    // joint 3: back torso (pivot spine)
    if (S0) w.twist(A_TH0 + 3, M.c1, spine);
    rot_y(M, TH(3));
    if (S0) w.twist(A_PHI3, M.c0, spine);
    rot_x(M, sn[A_PHI3], cs[A_PHI3]);
    if (S0) w.twist(A_PSI3, M.c2, spine);
    rot_z(M, sn[A_PSI3], cs[A_PSI3]);
    const Mat3T<V> M3 = M;
    const V tailb = fma3(-0.37f, M3.c0, spine);
    if (S0) w.marker(5, tailb);
    const V hip_mid = fma3(0.12f, M3.c0, fma3(-0.06f, M3.c2, tailb));
    const V lhip = fma3(0.08f, M3.c1, hip_mid);
    const V rhip = fma3(-0.08f, M3.c1, hip_mid);
    if (S2) w.marker(14, lhip);
    if (S2) w.marker(17, rhip);
    {
        Mat3T<V> Ml = M3;
        if (S2) w.twist(A_TH0 + 10, M3.c1, lhip);
        rot_y(Ml, TH(10));
        const V knee = fma3(-0.32f, Ml.c2, lhip);
        if (S2) w.marker(15, knee);
        if (S2) w.twist(A_TH0 + 11, M3.c1, knee);
        rot_y(Ml, TH(11));
        if (S2) w.marker(16, fma3(-0.25f, Ml.c2, knee));
        Mat3T<V> Mr = M3;
        if (S2) w.twist(A_TH0 + 12, M3.c1, rhip);
        rot_y(Mr, TH(12));
        const V kneer = fma3(-0.32f, Mr.c2, rhip);
        if (S2) w.marker(18, kneer);
        if (S2) w.twist(A_TH0 + 13, M3.c1, kneer);
        rot_y(Mr, TH(13));
        if (S2) w.marker(19, fma3(-0.25f, Mr.c2, kneer));
    }
    // joint 4: tail base (pivot tail_base), joint 5: tail mid (pivot tail1)
    if (S0) w.twist(A_TH0 + 4, M.c1, tailb);
    rot_y(M, TH(4));
    if (S0) w.twist(A_PSI4, M.c2, tailb);
    rot_z(M, sn[A_PSI4], cs[A_PSI4]);
    const V tail1 = fma3(-0.28f, M.c0, tailb);
    if (S0) w.marker(6, tail1);
    if (S0) w.twist(A_TH0 + 5, M.c1, tail1);
    rot_y(M, TH(5));
    if (S0) w.twist(A_PSI5, M.c2, tail1);
    rot_z(M, sn[A_PSI5], cs[A_PSI5]);
    if (S0) w.marker(7, fma3(-0.36f, M.c0, tail1));
#undef TH
}

// one thread per frame (fk_project, fte_jac)
__device__ __forceinline__ void cheetah_fk(const float2* __restrict__ sc, const FkWriter& w) {
    const Mat3 I = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    cheetah_fk_t<Col3, 7>(sc, I, w);
}
// thread i of a frame carries row i of the chain (fte_eval): writes component i of the markers and rotation axes
template <int SEG>
__device__ __forceinline__ void cheetah_fk_row(const float2* __restrict__ sc, const int i, float* p, float* tau) {
    const Mat3T<float> I = {i == 0 ? 1.f : 0.f, i == 1 ? 1.f : 0.f, i == 2 ? 1.f : 0.f};
    cheetah_fk_t<float, SEG>(sc, I, FkRowWriter{p + i, tau + i});
}

}  // namespace acino
