// Generic-skeleton FTE (the data-driven variant of the reference, /root/reference/src/build.py:28-302), fp64.
//
// The bodies below are written as phase loops `for (t = ctx.tid; t < n; t += ctx.nthreads)` separated by ctx.sync():
// on the GPU ctx is (threadIdx.x, blockDim.x, __syncthreads) - see skel.cu - and the very same source, compiled for
// the host with ctx = (0, 1, no-op), is what tests/host_harness runs on the CPU to check the kernels against the
// NumPy oracle before they ever see a GPU.  The harness is test infrastructure; the library has no host path.
//
// Forward kinematics of build.py:43-80 (quirks kept, see acinoset_b200/skeleton.py): a part's pose is the root
// (x, y, z) plus one increment per link on its path, d_l = M_a tv_l, where M_a is the LOCAL rotation
// Rz(psi_a) Rx(phi_a) Ry(theta_a) of the link's parent a (or its transpose, link flag) - rotations do not chain.
// So d pose_r / d angle_k(a) = sum over the links l of a on r's path of D_{l,k} = (d M_a / d angle_k) tv_l, and
//   g[a,k]          = sum_{l in links(a)} D_{l,k} . Sb_l            Sb_l      = sum_{r: l in path(r)} b_r
//   H[(a,k),(a',k')] = sum_{l,l'} D_{l,k}^T SA_{l,l'} D_{l',k'}      SA_{l,l'} = sum_{r: l,l' in path(r)} A_r
// with A_r, b_r the 3x3 normal block and gradient of output row r accumulated over the cameras.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SKEL_HD __host__ __device__ __forceinline__
#else
#define SKEL_HD inline
#endif

namespace acino {

constexpr int SK_MAX_PARTS = 32;
constexpr int SK_MAX_LINKS = 40;
constexpr int SK_MAX_CAMS = 16;

struct SkelCam {
    double R[9], t[3], fx, fy, cx, cy, D[4];
};

struct SkelDesc {
    int n_parts;                     // L: state is [x,y,z, phi_0..L-1, theta_0..L-1, psi_0..L-1], P = 3 + 3L
    int n_links;
    int n_out;                       // output rows (pose_dict order) = measured markers
    int n_cams;
    int dof_mask[SK_MAX_PARTS];      // bit0 phi (x), bit1 theta (y), bit2 psi (z)
    int link_parent[SK_MAX_LINKS];
    int link_flag[SK_MAX_LINKS];     // 1: parent's local rotation used transposed
    double link_tv[SK_MAX_LINKS][3];
    unsigned long long path[SK_MAX_PARTS];   // per output row: bit l set <=> link l is on its path
    int part_ptr[SK_MAX_PARTS + 1];  // CSR: links whose parent is part a
    int part_links[SK_MAX_LINKS];
    int loss_kind;                   // 0: redescending(a,b,c) of |w r| (all_optimizations.py:497); 1: |w r| (build.py:299)
    double la, lb, lc;               // redescending break points
    double delta;                    // loss_kind 1: curvature weight 1 / max(|w r|, delta)
    SkelCam cam[SK_MAX_CAMS];
};

// dynamic shared memory (doubles) of one frame
struct SkelSmemLayout {
    int d, D, pose, A, b, costp, Sb, SA, Atot, btot, total;
    SKEL_HD SkelSmemLayout(int n_links, int n_out) {
        int o = 0;
        d = o; o += n_links * 3;
        D = o; o += n_links * 9;
        pose = o; o += n_out * 3;
        A = o; o += n_out * 6;
        b = o; o += n_out * 3;
        costp = o; o += n_out;
        Sb = o; o += n_links * 3;
        SA = o; o += (n_links * (n_links + 1) / 2) * 6;
        Atot = o; o += 6;
        btot = o; o += 3;
        total = o;
    }
};

SKEL_HD int sk_pair_index(int l, int m, int n) {   // l <= m < n, row-major packed upper
    return l * n - (l * (l - 1)) / 2 + (m - l);
}

// literal logistic blend of build.py:382-395 and its derivative, e >= 0
SKEL_HD void sk_redescending(double a, double b, double c, double e, double& rho, double& drho, double& floor_) {
    const double sa = 1.0 / (1.0 + exp(-(e - a))), sb = 1.0 / (1.0 + exp(-(e - b))), sc = 1.0 / (1.0 + exp(-(e - c)));
    const double dsa = sa * (1 - sa), dsb = sb * (1 - sb), dsc = sc * (1 - sc);
    const double p1 = 0.5 * e * e, p2 = a * e - a * a / 2;
    const double u = (c - e) / (c - b), k3 = a * (c - b) / 2;
    const double p3 = a * b - a * a / 2 + k3 * (1 - u * u), dp3 = a * u;
    const double p4 = a * b - a * a / 2 + k3;
    rho = (1 - sa) * p1 + (sa - sb) * p2 + (sb - sc) * p3 + sc * p4;
    drho = (1 - sa) * e - dsa * p1 + (dsa - dsb) * p2 + (sa - sb) * a + (dsb - dsc) * p3 + (sb - sc) * dp3 + dsc * p4;
    floor_ = 1 - sa;
}

// c = a b (3x3 row-major)
SKEL_HD void sk_mm(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// One frame: x [P] -> cost, g [P], H [P(P+1)/2] (packed upper, row-major); any output may be NULL.
// meas [C][n_out][2], w [C][n_out].  sm: SkelSmemLayout(...).total doubles.
template <typename Ctx>
SKEL_HD void skel_eval_frame(const SkelDesc& S, const Ctx& ctx, const double* __restrict__ x, const double* __restrict__ meas,
                             const double* __restrict__ w, double* __restrict__ cost, double* __restrict__ g,
                             double* __restrict__ H, double* sm) {
    const int L = S.n_parts, NLk = S.n_links, NO = S.n_out, P = 3 + 3 * L;
    const SkelSmemLayout lay(NLk, NO);
    double* s_d = sm + lay.d;
    double* s_D = sm + lay.D;
    double* s_pose = sm + lay.pose;
    double* s_A = sm + lay.A;
    double* s_b = sm + lay.b;
    double* s_cost = sm + lay.costp;
    double* s_Sb = sm + lay.Sb;
    double* s_SA = sm + lay.SA;
    double* s_Atot = sm + lay.Atot;
    double* s_btot = sm + lay.btot;

    // ---- S1: link increments and their derivatives w.r.t. the parent's (phi, theta, psi)
    for (int l = ctx.tid; l < NLk; l += ctx.nthreads) {
        const int a = S.link_parent[l], m = S.dof_mask[a];
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double Ry[9], Rx[9], Rz[9], dRy[9], dRx[9], dRz[9];
        for (int i = 0; i < 9; ++i) { Ry[i] = Rx[i] = Rz[i] = I3[i]; dRy[i] = dRx[i] = dRz[i] = Z3[i]; }
        if (m & 2) {   // rot_y (build.py:407-414): [[c,0,-s],[0,1,0],[s,0,c]]
            const double c = cos(x[3 + L + a]), s = sin(x[3 + L + a]);
            Ry[0] = c; Ry[2] = -s; Ry[6] = s; Ry[8] = c;
            dRy[0] = -s; dRy[2] = -c; dRy[6] = c; dRy[8] = -s;
        }
        if (m & 1) {   // rot_x (:399-405): [[1,0,0],[0,c,s],[0,-s,c]]
            const double c = cos(x[3 + a]), s = sin(x[3 + a]);
            Rx[4] = c; Rx[5] = s; Rx[7] = -s; Rx[8] = c;
            dRx[4] = -s; dRx[5] = c; dRx[7] = -c; dRx[8] = -s;
        }
        if (m & 4) {   // rot_z (:416-423): [[c,s,0],[-s,c,0],[0,0,1]]
            const double c = cos(x[3 + 2 * L + a]), s = sin(x[3 + 2 * L + a]);
            Rz[0] = c; Rz[1] = s; Rz[3] = -s; Rz[4] = c;
            dRz[0] = -s; dRz[1] = c; dRz[3] = -c; dRz[4] = -s;
        }
        double XY[9], dXY_phi[9], dXY_th[9], M[4][9];
        sk_mm(Rx, Ry, XY);
        sk_mm(dRx, Ry, dXY_phi);
        sk_mm(Rx, dRy, dXY_th);
        sk_mm(Rz, XY, M[0]);          // M = Rz Rx Ry  (build.py:54-59)
        sk_mm(Rz, dXY_phi, M[1]);     // d/d phi
        sk_mm(Rz, dXY_th, M[2]);      // d/d theta
        sk_mm(dRz, XY, M[3]);         // d/d psi
        const double* tv = S.link_tv[l];
        const bool tr = S.link_flag[l] & 1;
        for (int q = 0; q < 4; ++q) {
            double* out = q == 0 ? s_d + 3 * l : s_D + 9 * l + 3 * (q - 1);
            for (int i = 0; i < 3; ++i)
                out[i] = tr ? (M[q][i] * tv[0] + M[q][3 + i] * tv[1] + M[q][6 + i] * tv[2])
                            : (M[q][3 * i] * tv[0] + M[q][3 * i + 1] * tv[1] + M[q][3 * i + 2] * tv[2]);
        }
    }
    ctx.sync();

    // ---- S2 + S3: pose of every output row, then projection + loss over the cameras
    for (int r = ctx.tid; r < NO; r += ctx.nthreads) {
        double p0 = x[0], p1 = x[1], p2 = x[2];
        const unsigned long long path = S.path[r];
        for (int l = 0; l < NLk; ++l)
            if ((path >> l) & 1ull) { p0 += s_d[3 * l]; p1 += s_d[3 * l + 1]; p2 += s_d[3 * l + 2]; }
        s_pose[3 * r] = p0; s_pose[3 * r + 1] = p1; s_pose[3 * r + 2] = p2;
        double A[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0}, cst = 0;
        for (int c = 0; c < S.n_cams; ++c) {
            const SkelCam& cam = S.cam[c];
            const double wt = w[c * NO + r];
            const double xc = cam.R[0] * p0 + cam.R[1] * p1 + cam.R[2] * p2 + cam.t[0];
            const double yc = cam.R[3] * p0 + cam.R[4] * p1 + cam.R[5] * p2 + cam.t[1];
            const double zc = cam.R[6] * p0 + cam.R[7] * p1 + cam.R[8] * p2 + cam.t[2];
            // pt3d_to_2d (build.py:457-473) and its Jacobian w.r.t. the camera-frame point
            const double iz = 1.0 / zc, aa = xc * iz, bb = yc * iz;
            const double r2 = aa * aa + bb * bb + 1e-12, rr = sqrt(r2), ir = 1.0 / rr;
            const double th = atan(rr), th2 = th * th;
            const double td = th * (1 + th2 * (cam.D[0] + th2 * (cam.D[1] + th2 * (cam.D[2] + th2 * cam.D[3]))));
            const double dtd = 1 + th2 * (3 * cam.D[0] + th2 * (5 * cam.D[1] + th2 * (7 * cam.D[2] + th2 * 9 * cam.D[3])));
            const double sd = td * ir;
            const double q = (dtd / (1 + r2) - sd) * (ir * ir);
            const double m00 = sd + aa * aa * q, m01 = aa * bb * q, m11 = sd + bb * bb * q;
            const double fxi = cam.fx * iz, fyi = cam.fy * iz;
            const double ju[3] = {fxi * m00, fxi * m01, -fxi * (m00 * aa + m01 * bb)};
            const double jv[3] = {fyi * m01, fyi * m11, -fyi * (m01 * aa + m11 * bb)};
            double Ju[3], Jv[3];      // world-frame rows: j^T R
            for (int j = 0; j < 3; ++j) {
                Ju[j] = ju[0] * cam.R[j] + ju[1] * cam.R[3 + j] + ju[2] * cam.R[6 + j];
                Jv[j] = jv[0] * cam.R[j] + jv[1] * cam.R[3 + j] + jv[2] * cam.R[6 + j];
            }
            const double res[2] = {wt != 0 ? cam.fx * aa * sd + cam.cx - meas[(c * NO + r) * 2] : 0.0,
                                   wt != 0 ? cam.fy * bb * sd + cam.cy - meas[(c * NO + r) * 2 + 1] : 0.0};
            for (int dd = 0; dd < 2; ++dd) {
                const double* J = dd ? Jv : Ju;
                const double e = fabs(wt * res[dd]);
                double rho, gw, hw;     // d rho / d r = gw w^2 r ; curvature weight hw w^2
                if (S.loss_kind == 0) {
                    double drho, fl;
                    sk_redescending(S.la, S.lb, S.lc, e, rho, drho, fl);
                    gw = e > 0 ? drho / e : 0.0;
                    hw = e > 0 ? fmax(gw, fl) : fl;
                } else {
                    rho = e;
                    gw = e > 0 ? 1.0 / e : 0.0;
                    hw = 1.0 / fmax(e, S.delta);
                }
                cst += rho;
                const double w2 = wt * wt, gs = gw * w2 * res[dd], hs = hw * w2;
                b[0] += gs * J[0]; b[1] += gs * J[1]; b[2] += gs * J[2];
                A[0] += hs * J[0] * J[0]; A[1] += hs * J[0] * J[1]; A[2] += hs * J[0] * J[2];
                A[3] += hs * J[1] * J[1]; A[4] += hs * J[1] * J[2]; A[5] += hs * J[2] * J[2];
            }
        }
        for (int i = 0; i < 6; ++i) s_A[6 * r + i] = A[i];
        for (int i = 0; i < 3; ++i) s_b[3 * r + i] = b[i];
        s_cost[r] = cst;
    }
    ctx.sync();

    // ---- S4: sums over the rows carried by each link / link pair (fixed order: rows ascending)
    const int n_pairs = NLk * (NLk + 1) / 2;
    for (int t = ctx.tid; t < n_pairs + NLk + 1; t += ctx.nthreads) {
        if (t < n_pairs) {
            int l = 0, rem = t;
            while (rem >= NLk - l) { rem -= NLk - l; ++l; }
            const int m = l + rem;
            const unsigned long long need = (1ull << l) | (1ull << m);
            double a[6] = {0, 0, 0, 0, 0, 0};
            for (int r = 0; r < NO; ++r)
                if ((S.path[r] & need) == need)
                    for (int i = 0; i < 6; ++i) a[i] += s_A[6 * r + i];
            for (int i = 0; i < 6; ++i) s_SA[6 * t + i] = a[i];
        } else if (t < n_pairs + NLk) {
            const int l = t - n_pairs;
            double a[3] = {0, 0, 0};
            for (int r = 0; r < NO; ++r)
                if ((S.path[r] >> l) & 1ull)
                    for (int i = 0; i < 3; ++i) a[i] += s_b[3 * r + i];
            for (int i = 0; i < 3; ++i) s_Sb[3 * l + i] = a[i];
        } else {
            double a[6] = {0, 0, 0, 0, 0, 0}, bt[3] = {0, 0, 0}, c = 0;
            for (int r = 0; r < NO; ++r) {
                for (int i = 0; i < 6; ++i) a[i] += s_A[6 * r + i];
                for (int i = 0; i < 3; ++i) bt[i] += s_b[3 * r + i];
                c += s_cost[r];
            }
            for (int i = 0; i < 6; ++i) s_Atot[i] = a[i];
            for (int i = 0; i < 3; ++i) s_btot[i] = bt[i];
            if (cost) *cost = c;
        }
    }
    ctx.sync();

    // ---- S5: gradient and packed upper triangle.  slot p >= 3: k = (p-3) / L (0 phi, 1 theta, 2 psi), a = (p-3) % L
    if (g)
        for (int p = ctx.tid; p < P; p += ctx.nthreads) {
            if (p < 3) { g[p] = s_btot[p]; continue; }
            const int k = (p - 3) / L, a = (p - 3) - k * L;
            double acc = 0;
            for (int i = S.part_ptr[a]; i < S.part_ptr[a + 1]; ++i) {
                const int l = S.part_links[i];
                const double* Dk = s_D + 9 * l + 3 * k;
                acc += Dk[0] * s_Sb[3 * l] + Dk[1] * s_Sb[3 * l + 1] + Dk[2] * s_Sb[3 * l + 2];
            }
            g[p] = acc;
        }
    if (H) {
        const int NU = P * (P + 1) / 2;
        for (int idx = ctx.tid; idx < NU; idx += ctx.nthreads) {
            int p = 0, rem = idx;
            while (rem >= P - p) { rem -= P - p; ++p; }
            const int q = p + rem;
            double acc = 0;
            if (q < 3) {
                const int ii = p == 0 ? q : (p == 1 ? 2 + q : 5);     // (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
                acc = s_Atot[ii];
            } else {
                const int kq = (q - 3) / L, aq = (q - 3) - kq * L;
                for (int i = S.part_ptr[aq]; i < S.part_ptr[aq + 1]; ++i) {
                    const int lq = S.part_links[i];
                    const double* Dq = s_D + 9 * lq + 3 * kq;
                    if (p < 3) {
                        const double* a = s_SA + 6 * sk_pair_index(lq, lq, NLk);
                        const double row[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
                        acc += row[p][0] * Dq[0] + row[p][1] * Dq[1] + row[p][2] * Dq[2];
                    } else {
                        const int kp = (p - 3) / L, ap = (p - 3) - kp * L;
                        for (int j = S.part_ptr[ap]; j < S.part_ptr[ap + 1]; ++j) {
                            const int lp = S.part_links[j];
                            const double* Dp = s_D + 9 * lp + 3 * kp;
                            const double* a = s_SA + 6 * (lp <= lq ? sk_pair_index(lp, lq, NLk) : sk_pair_index(lq, lp, NLk));
                            const double y0 = a[0] * Dq[0] + a[1] * Dq[1] + a[2] * Dq[2];
                            const double y1 = a[1] * Dq[0] + a[3] * Dq[1] + a[4] * Dq[2];
                            const double y2 = a[2] * Dq[0] + a[4] * Dq[1] + a[5] * Dq[2];
                            acc += Dp[0] * y0 + Dp[1] * y1 + Dp[2] * y2;
                        }
                    }
                }
            }
            H[idx] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Levenberg-Marquardt building blocks for the generic variant.  Objective (build.py:287-302 with the dynamics of
// :231-261 eliminated exactly like SURVEY.md B6): F = sum rho(w r) + sum_{n>=3,p} q_p (third difference / h^2)^2,
// sw[p] = 2 q_p / h^4; box bounds lo/hi per slot (build.py:263-266), `last_free` = the reference's range(1, N) leaves
// the last frame unbounded.

SKEL_HD double sk_d3tsd3(const double* x, int N, int P, int n, int p) {   // (D3^T D3 x)[n][p]
    const double c[4] = {1, -3, 3, -1};
    double acc = 0;
    for (int m = n > 3 ? n : 3; m <= n + 3 && m < N; ++m) {       // rows m use frames m..m-3, coefficient c[m - frame]
        const double tdiff = x[(size_t)m * P + p] - 3 * x[(size_t)(m - 1) * P + p] + 3 * x[(size_t)(m - 2) * P + p] - x[(size_t)(m - 3) * P + p];
        acc += c[m - n] * tdiff;
    }
    return acc;
}

SKEL_HD double sk_smooth_entry(int N, int n, int k) {   // (D3^T D3)[n][n-k], k = 0..3
    const double c[4] = {1, -3, 3, -1};
    double acc = 0;
    for (int m = n > 3 ? n : 3; m <= n - k + 3 && m < N; ++m) acc += c[m - n] * c[m - (n - k)];
    return acc;
}

// per (frame, slot): total gradient and the frozen flag; per frame: smoothness cost
template <typename Ctx>
SKEL_HD void skel_prepare(const Ctx& ctx, int N, int P, int last_free, const double* x, const double* g, const double* sw,
                          const double* lo, const double* hi, double* gtot, unsigned char* fixed, double* cost_s) {
    for (long long i = ctx.tid; i < (long long)N * P; i += ctx.nthreads) {
        const int n = (int)(i / P), p = (int)(i - (long long)n * P);
        const double gt = g[i] + sw[p] * sk_d3tsd3(x, N, P, n, p);
        gtot[i] = gt;
        const bool bounded = !(last_free && n == N - 1);
        fixed[i] = bounded && ((x[i] <= lo[p] && gt > 0) || (x[i] >= hi[p] && gt < 0));
    }
    for (int n = ctx.tid; n < N; n += ctx.nthreads) {
        double c = 0;
        if (n >= 3)
            for (int p = 0; p < P; ++p) {
                const double td = x[(size_t)n * P + p] - 3 * x[(size_t)(n - 1) * P + p] + 3 * x[(size_t)(n - 2) * P + p] - x[(size_t)(n - 3) * P + p];
                c += 0.5 * sw[p] * td * td;
            }
        cost_s[n] = c;
    }
}

// lower band of (B + lam diag B) with frozen rows/columns replaced by identity; AB [N P][3P + 1], AB[i][k] = B[i][i-k]
template <typename Ctx>
SKEL_HD void skel_assemble(const Ctx& ctx, int N, int P, const double* H, const double* gtot, const unsigned char* fixed,
                           const double* sw, double lam, double* AB, double* rhs) {
    const int hb = 3 * P, W = hb + 1;
    const long long n_rows = (long long)N * P;
    for (long long e = ctx.tid; e < n_rows * W; e += ctx.nthreads) {
        const long long i = e / W;
        const int k = (int)(e - i * W);
        const long long j = i - k;
        double v = 0;
        if (j >= 0) {
            const int n = (int)(i / P), p = (int)(i - (long long)n * P);
            const int nj = (int)(j / P), pj = (int)(j - (long long)nj * P);
            if (nj == n) v = H[(size_t)n * (P * (P + 1) / 2) + (pj * P - (pj * (pj - 1)) / 2 + (p - pj))];
            if (pj == p) v += sw[p] * sk_smooth_entry(N, n, n - nj);
            if (k == 0) v *= 1 + lam;
            if (fixed[i] || fixed[j]) v = k == 0 ? 1.0 : 0.0;
        }
        AB[e] = v;
    }
    for (long long i = ctx.tid; i < n_rows; i += ctx.nthreads) rhs[i] = fixed[i] ? 0.0 : -gtot[i];
}

// In-place band Cholesky B = L L^T (row-wise lower band storage AB[i][k] = B[i][i-k], half bandwidth hb) and solve of
// B x = rhs, by ONE cooperating group of threads (a CTA).  Blocked right-looking, panel width NB (compile time):
//   A   the panel (rows j0 .. j0+NB-1+hb, NB columns) and its right-hand side entries are staged in shared memory;
//   B1  its NB x NB triangle (+ the right-hand side as an extra row: forward substitution for free) is factored column
//       by column - one barrier per column, a handful of warps busy, the pivot's 1/p and 1/sqrt(p) computed once by
//       the thread that finishes the pivot (Newton on the fp32 rsqrt) and handed over through a double-buffered slot;
//   B2  the hb rows below the triangle are a triangular solve with NB columns per row: one thread per row, the row in
//       registers, no barriers;
//   C   write-back, then ONE rank-NB update of the trailing band per panel, register tiled (4 x 4 strided rows), so
//       every entry of the band makes one global round trip per panel instead of one per column.
// The backward substitution walks the panels right to left.  info = index + 1 of the first non-positive pivot.
// sm: band_panel_doubles(hb, NB) doubles.  (ncu of the first, column-at-a-time-over-the-whole-panel version: 47 % of
// the instructions in the panel step, 45 % in the trailing update, 32 warps waiting on one - profiles/r01_skel.md.)
SKEL_HD size_t band_panel_doubles(int hb, int nb) {
    const size_t fwd = (size_t)(nb + hb) * (nb + 1), bwd = (size_t)nb * 32 + (size_t)nb * nb;
    return (fwd > bwd ? fwd : bwd) + 2 * nb + 8;
}

SKEL_HD double sk_rsqrt(double a) {
#ifdef __CUDA_ARCH__
    if (a > 1e-30 && a < 1e30) {
        double y = (double)rsqrtf((float)a);           // 22 bits, then three Newton steps
        y = y * (1.5 - 0.5 * a * y * y);
        y = y * (1.5 - 0.5 * a * y * y);
        y = y * (1.5 - 0.5 * a * y * y);
        return y;
    }
#endif
    return 1.0 / sqrt(a);
}

template <int NB, typename Ctx>
SKEL_HD void band_cholesky_solve(const Ctx& ctx, long long n, int hb, double* AB, double* x, int* info, double* sm) {
    constexpr int LD = NB + 1;
    constexpr int NBQ = NB <= 4 ? 4 : (NB <= 8 ? 8 : (NB <= 16 ? 16 : 32));   // power of two >= NB (NB <= 32)
    const int W = hb + 1;
    double* Lp = sm;                                   // [(NB + hb)][LD] panel rows, column c at Lp[r * LD + c]
    double* yp = sm + band_panel_doubles(hb, NB) - 2 * NB - 8;   // [NB] right-hand side of the panel columns
    double* isd = yp + NB;                             // [NB] 1 / L_cc
    double* scal = isd + NB;                           // 2 x {pivot, 1 / pivot, 1 / sqrt(pivot)}, [6] = failure flag
    if (ctx.tid == 0) scal[6] = 0.0;
    const int tx = ctx.tid % NBQ, ty = ctx.tid / NBQ;
    const int sx = ctx.nthreads < NBQ ? ctx.nthreads : NBQ, sy = ctx.nthreads / NBQ > 0 ? ctx.nthreads / NBQ : 1;
    for (long long j0 = 0; j0 < n; j0 += NB) {
        const int nbp = (int)((n - j0) < NB ? (n - j0) : NB);
        const long long j1 = j0 + nbp;
        const long long rmax = (j1 - 1 + hb) < (n - 1) ? (j1 - 1 + hb) : (n - 1);
        const int Rn = (int)(rmax - j0 + 1);           // panel rows
        // ---- A: stage the panel
        for (int r = ty; r < Rn; r += sy)
            for (int c = tx; c < nbp; c += sx) {
                const int k = r - c;                   // distance below the diagonal of column j0 + c
                Lp[r * LD + c] = (k >= 0 && k <= hb) ? AB[(j0 + r) * W + k] : 0.0;
            }
        for (int c = ctx.tid; c < nbp; c += ctx.nthreads) yp[c] = x[j0 + c];
        if (ctx.tid == 0) {
            const double piv = AB[j0 * W];
            scal[0] = piv;
            const double y = piv > 0 ? sk_rsqrt(piv) : 0.0;
            scal[1] = y * y;
            scal[2] = y;
        }
        ctx.sync();
        // ---- B1: the nbp x nbp triangle and the right-hand side row (row index nbp here).  Only the first NPART threads
        //      take part and synchronise among themselves (named barrier): 23 idle warps arriving at 16 CTA-wide
        //      barriers per panel cost more than the triangle itself (ncu: 26 % of the kernel waiting there)
        constexpr int NPART = ((NBQ * (NB + 1) + 31) / 32) * 32 < 1024 ? ((NBQ * (NB + 1) + 31) / 32) * 32 : 1024;   // the kernel runs 1024 threads
        if (ctx.tid < NPART) {
            double inv_prev = 0;
            for (int c = 0; c < nbp; ++c) {
                const double* sc = scal + 3 * (c & 1);
                double* sn = scal + 3 * ((c + 1) & 1);
                const double piv = sc[0], ip = sc[1];
                if (!(piv > 0)) {                      // uniform among the participants: all read the same pivot
                    if (ctx.tid == 0) {
                        if (*info == 0) *info = (int)(j0 + c + 1);
                        scal[6] = 1.0;
                    }
                    break;
                }
                const int nc = nbp - 1 - c;            // columns right of c
                for (int rr = ty; rr <= nc; rr += sy) {    // triangle rows c+1 .. nbp-1, then the right-hand side (rr == nc)
                    const bool rhs = rr == nc;
                    const int r = c + 1 + rr;
                    const double lrc = rhs ? yp[c] : Lp[r * LD + c];
                    for (int q = tx; q < nc; q += sx) {
                        const int cc = c + 1 + q;
                        const double m = Lp[cc * LD + c] * ip;
                        if (rhs) {
                            yp[cc] -= lrc * m;
                        } else if (r >= cc) {
                            const double v = Lp[r * LD + cc] - lrc * m;
                            Lp[r * LD + cc] = v;
                            if (r == cc && q == 0) {   // the next pivot is final: prepare its scalars
                                sn[0] = v;
                                const double y = v > 0 ? sk_rsqrt(v) : 0.0;
                                sn[1] = y * y;
                                sn[2] = y;
                            }
                        }
                    }
                }
                if (c > 0)
                    for (int r = c - 1 + ctx.tid; r <= nbp; r += ctx.nthreads) {   // scale column c-1 (nobody reads it any more)
                        if (r < nbp) Lp[r * LD + c - 1] *= inv_prev;
                        else yp[c - 1] *= inv_prev;
                    }
                inv_prev = sc[2];
                if (ctx.tid == 0) isd[c] = inv_prev;
                ctx.sync_part(NPART);
            }
            for (int r = nbp - 1 + ctx.tid; r <= nbp; r += ctx.nthreads) {         // the last column
                if (r < nbp) Lp[r * LD + nbp - 1] *= inv_prev;
                else yp[nbp - 1] *= inv_prev;
            }
        }
        ctx.sync();
        if (scal[6] != 0.0) return;                    // non-positive pivot (info is set)
        // ---- B2: rows below the triangle: l[c] = (a[c] - sum_{k<c} l[k] L11[c][k]) / L11[c][c], one thread per row
        for (int r = nbp + ctx.tid; r < Rn; r += ctx.nthreads) {
            double l[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) l[c] = c < nbp ? Lp[r * LD + c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < nbp) {
                    double v = l[c];
#pragma unroll
                    for (int k = 0; k < c; ++k) v -= l[k] * Lp[c * LD + k];
                    l[c] = v * isd[c];
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nbp) Lp[r * LD + c] = l[c];
        }
        ctx.sync();
        // ---- C: write the panel back; rank-nbp update of the trailing band and of the right-hand side
        for (int r = ty; r < Rn; r += sy)
            for (int c = tx; c < nbp; c += sx) {
                const int k = r - c;
                if (k >= 0 && k <= hb) AB[(j0 + r) * W + k] = Lp[r * LD + c];
            }
        for (int c = ctx.tid; c < nbp; c += ctx.nthreads) x[j0 + c] = yp[c];
        const int T = (int)(rmax - j1 + 1);            // trailing rows j1 .. rmax
        // register-tiled update: a task owns rows {ta + i nt} x {tb + j nt}, i, j < 4 (strided so that neighbouring
        // threads read neighbouring panel rows: conflict-free with the odd row stride LD); only the entries on or below
        // the diagonal (i > j, or i == j with tb <= ta) exist
        const int nt = (T + 3) / 4;
        for (int e = ctx.tid; e < nt * nt + T; e += ctx.nthreads) {
            if (e >= nt * nt) {                        // the right-hand side row
                const int a = e - nt * nt;
                const double* La = Lp + (size_t)(nbp + a) * LD;
                double acc = 0;
                for (int c = 0; c < NB; ++c) acc += c < nbp ? La[c] * yp[c] : 0.0;
                x[j1 + a] -= acc;
                continue;
            }
            const int ta = e / nt, tb = e - ta * nt;
            const double* Pa[4];
            const double* Pb[4];
            for (int i = 0; i < 4; ++i) {              // rows past the end read row 0 of the trailing part; never stored
                const int a = ta + i * nt, b = tb + i * nt;
                Pa[i] = Lp + (size_t)(nbp + (a < T ? a : 0)) * LD;
                Pb[i] = Lp + (size_t)(nbp + (b < T ? b : 0)) * LD;
            }
            double acc[4][4];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) acc[i][j] = 0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < nbp) {
                    double la[4], lb[4];
                    for (int i = 0; i < 4; ++i) { la[i] = Pa[i][c]; lb[i] = Pb[i][c]; }
                    for (int i = 0; i < 4; ++i)
                        for (int j = 0; j <= i; ++j) acc[i][j] += la[i] * lb[j];
                }
            }
            // read-modify-write of the band: all loads first, then all stores (a store followed by a load of the same
            // array is kept in order by the compiler: one L2 round trip per entry otherwise)
            double old[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const int a = ta + i * nt, b = tb + j * nt;
                    old[i][j] = (a < T && b <= a) ? AB[(j1 + a) * W + (a - b)] : 0.0;
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const int a = ta + i * nt, b = tb + j * nt;
                    if (a < T && b <= a) AB[(j1 + a) * W + (a - b)] = old[i][j] - acc[i][j];
                }
        }
        ctx.sync();
    }
    // ---- backward substitution L^T x = y, panels right to left.  Per panel: the contributions of the already
    //      solved rows in NSEG interleaved partial sums per column (fixed order), the triangle from shared memory
    constexpr int NSEG = 32;
    double* part = sm;                                 // [NB][NSEG]
    double* tri = sm + (size_t)NB * NSEG;              // [NB][NB]
    const long long n_panels = (n + NB - 1) / NB;
    for (long long p = n_panels - 1; p >= 0; --p) {
        const long long j0 = p * NB;
        const int nbp = (int)((n - j0) < NB ? (n - j0) : NB);
        const long long j1 = j0 + nbp;
        for (int e = ctx.tid; e < nbp * NSEG + nbp * NB; e += ctx.nthreads) {
            if (e < nbp * NSEG) {
                const int c = e / NSEG, sg = e - c * NSEG;
                const long long j = j0 + c;
                const long long imax = (j + hb) < (n - 1) ? (j + hb) : (n - 1);
                double acc = 0;
                for (long long i = j1 + sg; i <= imax; i += NSEG) acc += AB[i * W + (i - j)] * x[i];
                part[e] = acc;
            } else {
                const int q = e - nbp * NSEG, r = q / NB, c = q - r * NB;
                tri[r * NB + c] = (c < nbp && r >= c && r - c <= hb) ? AB[(j0 + r) * W + (r - c)] : 0.0;
            }
        }
        for (int c = ctx.tid; c < nbp; c += ctx.nthreads) yp[c] = x[j0 + c];
        ctx.sync();
        for (int c = ctx.tid; c < nbp; c += ctx.nthreads) {       // fold the partial sums (fixed order)
            double sacc = yp[c];
            for (int sg = 0; sg < NSEG; ++sg) sacc -= part[c * NSEG + sg];
            yp[c] = sacc;
        }
        ctx.sync();
        if (ctx.tid == 0) {
            double xs[NB];
#pragma unroll
            for (int c = NB - 1; c >= 0; --c) {
                xs[c] = 0;
                if (c < nbp) {
                    double sacc = yp[c];
#pragma unroll
                    for (int r = c + 1; r < NB; ++r) sacc -= r < nbp ? tri[r * NB + c] * xs[r] : 0.0;
                    xs[c] = sacc / tri[c * NB + c];
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nbp) x[j0 + c] = xs[c];
        }
        ctx.sync();
    }
}

// trial point xt = clip(x + d)
template <typename Ctx>
SKEL_HD void skel_trial(const Ctx& ctx, int N, int P, int last_free, const double* x, const double* d, const double* lo,
                        const double* hi, double* xt) {
    for (long long i = ctx.tid; i < (long long)N * P; i += ctx.nthreads) {
        const int n = (int)(i / P), p = (int)(i - (long long)n * P);
        double v = x[i] + d[i];
        if (!(last_free && n == N - 1)) v = fmin(fmax(v, lo[p]), hi[p]);
        xt[i] = v;
    }
}

// quadratic-model reduction and step norm per frame, s = xt - x:
//   pred[n] = -(gtot . s) - 1/2 s^T H_n s - 1/2 sw (D3 s)_n^2, step[n] = max |s|
template <typename Ctx>
SKEL_HD void skel_pred(const Ctx& ctx, int N, int P, const double* x, const double* xt, const double* gtot, const double* H,
                       const double* sw, double* pred, double* step) {
    for (int n = ctx.tid; n < N; n += ctx.nthreads) {
        const double* Hn = H + (size_t)n * (P * (P + 1) / 2);
        double lin = 0, quad = 0, smax = 0;
        for (int p = 0; p < P; ++p) {
            const double sp = xt[(size_t)n * P + p] - x[(size_t)n * P + p];
            lin += gtot[(size_t)n * P + p] * sp;
            smax = fmax(smax, fabs(sp));
            double row = 0.5 * Hn[p * P - (p * (p - 1)) / 2] * sp;
            for (int q = p + 1; q < P; ++q) row += Hn[p * P - (p * (p - 1)) / 2 + (q - p)] * (xt[(size_t)n * P + q] - x[(size_t)n * P + q]);
            quad += sp * row;
            if (n >= 3) {
                const double td = (xt[(size_t)n * P + p] - x[(size_t)n * P + p]) - 3 * (xt[(size_t)(n - 1) * P + p] - x[(size_t)(n - 1) * P + p]) +
                                  3 * (xt[(size_t)(n - 2) * P + p] - x[(size_t)(n - 2) * P + p]) - (xt[(size_t)(n - 3) * P + p] - x[(size_t)(n - 3) * P + p]);
                quad += 0.5 * sw[p] * td * td;
            }
        }
        pred[n] = -lin - quad;
        step[n] = smax;
    }
}

}  // namespace acino
