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
__device__ __forceinline__ Col3 cross(Col3 a, Col3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
struct Mat3 {  // body->world rotation R_k_I stored by columns
    Col3 c0, c1, c2;
};
// M <- M * Ry_a(th):  (c0,c2) <- (c c0 - s c2, s c0 + c c2)       [rot_y transposed, :75-82]
__device__ __forceinline__ void rot_y(Mat3& M, float s, float c) {
    const Col3 a = M.c0, b = M.c2;
    M.c0 = fma3(c, a, (-s) * b);
    M.c2 = fma3(s, a, c * b);
}
// M <- M * Rx_a(ph):  (c1,c2) <- (c c1 + s c2, -s c1 + c c2)      [rot_x transposed, :66-73]
__device__ __forceinline__ void rot_x(Mat3& M, float s, float c) {
    const Col3 a = M.c1, b = M.c2;
    M.c1 = fma3(c, a, s * b);
    M.c2 = fma3(-s, a, c * b);
}
// M <- M * Rz_a(ps):  (c0,c1) <- (c c0 + s c1, -s c0 + c c1)      [rot_z transposed, :84-91]
__device__ __forceinline__ void rot_z(Mat3& M, float s, float c) {
    const Col3 a = M.c0, b = M.c1;
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

// angle slot ids
enum { A_PHI0 = 0, A_PHI1 = 1, A_PHI3 = 2, A_TH0 = 3, A_PSI0 = 17, A_PSI1 = 18, A_PSI3 = 19, A_PSI4 = 20, A_PSI5 = 21 };

// Cheetah forward kinematics of one frame: marker positions relative to the head point and the
// world-frame twist of every angle.  Follows the chain RI_0..RI_13 (:101-128) and p_* (:138-165);
// R_k_I = R_parent_I Ry_a(theta) Rx_a(phi) Rz_a(psi), and the world axis of each angle is the
// matching column of the partially composed matrix (theta: column 1 before Ry; phi: column 0
// after Ry; psi: column 2 after Rx).
__device__ __forceinline__ void cheetah_fk(const float2* __restrict__ sc, const FkWriter& w) {
    float sn[NANG], cs[NANG];
#pragma unroll
    for (int i = 0; i < NANG; ++i) {
        const float2 v = sc[i];
        sn[i] = v.x;
        cs[i] = v.y;
    }
#define TH(k) sn[A_TH0 + (k)], cs[A_TH0 + (k)]
    const Col3 zero = {0.f, 0.f, 0.f};
    // joint 0: head
    Mat3 M = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    w.twist0(A_TH0 + 0, M.c1);
    rot_y(M, TH(0));
    w.twist0(A_PHI0, M.c0);
    rot_x(M, sn[A_PHI0], cs[A_PHI0]);
    w.twist0(A_PSI0, M.c2);
    rot_z(M, sn[A_PSI0], cs[A_PSI0]);
    w.marker(0, 0.03f * M.c1);                          // l_eye
    w.marker(1, -0.03f * M.c1);                         // r_eye
    w.marker(2, 0.055f * (M.c0 - M.c2));                // nose
    // joint 1: neck
    w.twist0(A_TH0 + 1, M.c1);
    rot_y(M, TH(1));
    w.twist0(A_PHI1, M.c0);
    rot_x(M, sn[A_PHI1], cs[A_PHI1]);
    w.twist0(A_PSI1, M.c2);
    rot_z(M, sn[A_PSI1], cs[A_PSI1]);
    const Col3 neck = -0.28f * M.c0;
    w.marker(3, neck);
    // joint 2: front torso (pivot neck_base)
    w.twist(A_TH0 + 2, M.c1, neck);
    rot_y(M, TH(2));
    const Mat3 M2 = M;
    const Col3 spine = fma3(-0.37f, M2.c0, neck);
    w.marker(4, spine);
    const Col3 sh_mid = fma3(-0.04f, M2.c0, fma3(-0.10f, M2.c2, neck));
    const Col3 lsh = fma3(0.08f, M2.c1, sh_mid);
    const Col3 rsh = fma3(-0.08f, M2.c1, sh_mid);
    w.marker(8, lsh);
    w.marker(11, rsh);
    // joints 6,7: left front leg; 8,9: right front leg (axis = column 1 of M2 throughout)
    {
        Mat3 Ml = M2;
        w.twist(A_TH0 + 6, M2.c1, lsh);
        rot_y(Ml, TH(6));
        const Col3 knee = fma3(-0.24f, Ml.c2, lsh);
        w.marker(9, knee);
        w.twist(A_TH0 + 7, M2.c1, knee);
        rot_y(Ml, TH(7));
        w.marker(10, fma3(-0.28f, Ml.c2, knee));
        Mat3 Mr = M2;
        w.twist(A_TH0 + 8, M2.c1, rsh);
        rot_y(Mr, TH(8));
        const Col3 kneer = fma3(-0.24f, Mr.c2, rsh);
        w.marker(12, kneer);
        w.twist(A_TH0 + 9, M2.c1, kneer);
        rot_y(Mr, TH(9));
        w.marker(13, fma3(-0.28f, Mr.c2, kneer));
    }
    // joint 3: back torso (pivot spine)
    w.twist(A_TH0 + 3, M.c1, spine);
    rot_y(M, TH(3));
    w.twist(A_PHI3, M.c0, spine);
    rot_x(M, sn[A_PHI3], cs[A_PHI3]);
    w.twist(A_PSI3, M.c2, spine);
    rot_z(M, sn[A_PSI3], cs[A_PSI3]);
    const Mat3 M3 = M;
    const Col3 tailb = fma3(-0.37f, M3.c0, spine);
    w.marker(5, tailb);
    const Col3 hip_mid = fma3(0.12f, M3.c0, fma3(-0.06f, M3.c2, tailb));
    const Col3 lhip = fma3(0.08f, M3.c1, hip_mid);
    const Col3 rhip = fma3(-0.08f, M3.c1, hip_mid);
    w.marker(14, lhip);
    w.marker(17, rhip);
    {
        Mat3 Ml = M3;
        w.twist(A_TH0 + 10, M3.c1, lhip);
        rot_y(Ml, TH(10));
        const Col3 knee = fma3(-0.32f, Ml.c2, lhip);
        w.marker(15, knee);
        w.twist(A_TH0 + 11, M3.c1, knee);
        rot_y(Ml, TH(11));
        w.marker(16, fma3(-0.25f, Ml.c2, knee));
        Mat3 Mr = M3;
        w.twist(A_TH0 + 12, M3.c1, rhip);
        rot_y(Mr, TH(12));
        const Col3 kneer = fma3(-0.32f, Mr.c2, rhip);
        w.marker(18, kneer);
        w.twist(A_TH0 + 13, M3.c1, kneer);
        rot_y(Mr, TH(13));
        w.marker(19, fma3(-0.25f, Mr.c2, kneer));
    }
    // joint 4: tail base (pivot tail_base), joint 5: tail mid (pivot tail1)
    w.twist(A_TH0 + 4, M.c1, tailb);
    rot_y(M, TH(4));
    w.twist(A_PSI4, M.c2, tailb);
    rot_z(M, sn[A_PSI4], cs[A_PSI4]);
    const Col3 tail1 = fma3(-0.28f, M.c0, tailb);
    w.marker(6, tail1);
    w.twist(A_TH0 + 5, M.c1, tail1);
    rot_y(M, TH(5));
    w.twist(A_PSI5, M.c2, tail1);
    rot_z(M, sn[A_PSI5], cs[A_PSI5]);
    w.marker(7, fma3(-0.36f, M.c0, tail1));
    (void)zero;
#undef TH
}

}  // namespace acino
