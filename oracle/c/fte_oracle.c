/*
 * ORACLE - TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Plain-C fp64 restatement of the reference's FTE measurement term for the cheetah skeleton:
 *   rot_x/rot_y/rot_z + chain RI_0..RI_13    /root/reference/src/all_optimizations.py:66-128
 *   marker positions                          :138-179
 *   pt3d_to_2d                                :193-209
 *   redescending_loss                         /root/reference/src/build.py:382-395
 *   measurement residual / weights / objective all_optimizations.py:302-308,394-399,494-497
 * with the analytic derivatives IPOPT would obtain by AD.  It is validated against the NumPy
 * oracle (tests/test_oracle_c.py) which is itself pinned to the reference's golden vectors.
 * Used only as the checker in tests and as the CPU baseline / --impl reference arm of bench.py
 * (kind "port": the reference's own Pyomo+IPOPT path cannot run in this image).
 * Deliberately straightforward: dense 2x25 Jacobian rows per observation, no algebraic tricks.
 */
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NA 25
#define NL 20
#define NJ 14

typedef struct { double m[3][3]; } M3;

static M3 rot_x(double a) { double c = cos(a), s = sin(a); M3 r = {{{1, 0, 0}, {0, c, s}, {0, -s, c}}}; return r; }
static M3 rot_y(double a) { double c = cos(a), s = sin(a); M3 r = {{{c, 0, -s}, {0, 1, 0}, {s, 0, c}}}; return r; }
static M3 rot_z(double a) { double c = cos(a), s = sin(a); M3 r = {{{c, s, 0}, {-s, c, 0}, {0, 0, 1}}}; return r; }
static M3 mul(M3 a, M3 b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
/* v_out = R^T v  (R_k_I = RI_k^T) */
static void tmulv(const M3* R, const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = R->m[0][i] * v[0] + R->m[1][i] * v[1] + R->m[2][i] * v[2];
}
static void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* active slot order: x,y,z, phi0,phi1,phi3, theta0..13, psi0,psi1,psi3,psi4,psi5 */
enum { S_PHI0 = 3, S_PHI1 = 4, S_PHI3 = 5, S_TH0 = 6, S_PSI0 = 20, S_PSI1 = 21, S_PSI3 = 22, S_PSI4 = 23, S_PSI5 = 24 };

static const int joint_parent[NJ] = {-1, 0, 1, 2, 3, 4, 2, 6, 2, 8, 3, 10, 3, 12};
/* segments: marker, parent marker (-1 = head point), joint, offset */
static const struct { int parent; int joint; double off[3]; } seg[NL] = {
    {-1, 0, {0, 0.03, 0}}, {-1, 0, {0, -0.03, 0}}, {-1, 0, {0.055, 0, -0.055}},
    {-1, 1, {-0.28, 0, 0}}, {3, 2, {-0.37, 0, 0}}, {4, 3, {-0.37, 0, 0}},
    {5, 4, {-0.28, 0, 0}}, {6, 5, {-0.36, 0, 0}},
    {3, 2, {-0.04, 0.08, -0.10}}, {8, 6, {0, 0, -0.24}}, {9, 7, {0, 0, -0.28}},
    {3, 2, {-0.04, -0.08, -0.10}}, {11, 8, {0, 0, -0.24}}, {12, 9, {0, 0, -0.28}},
    {5, 3, {0.12, 0.08, -0.06}}, {14, 10, {0, 0, -0.32}}, {15, 11, {0, 0, -0.25}},
    {5, 3, {0.12, -0.08, -0.06}}, {17, 12, {0, 0, -0.32}}, {18, 13, {0, 0, -0.25}},
};
/* pivot marker of each joint (-1 = head point) */
static const int joint_pivot[NJ] = {-1, -1, 3, 4, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18};

static int is_ancestor(int a, int k) { /* a ancestor-or-self of k */
    while (k >= 0) { if (k == a) return 1; k = joint_parent[k]; }
    return 0;
}

/* positions P[20][3] and Jacobian Jfk[20][3][25] */
static void cheetah_fk(const double* x, double P[NL][3], double Jfk[NL][3][NA]) {
    double phi[NJ] = {0}, th[NJ], psi[NJ] = {0};
    phi[0] = x[S_PHI0]; phi[1] = x[S_PHI1]; phi[3] = x[S_PHI3];
    for (int k = 0; k < NJ; ++k) th[k] = x[S_TH0 + k];
    psi[0] = x[S_PSI0]; psi[1] = x[S_PSI1]; psi[3] = x[S_PSI3]; psi[4] = x[S_PSI4]; psi[5] = x[S_PSI5];
    M3 RI[NJ];
    RI[0] = mul(rot_z(psi[0]), mul(rot_x(phi[0]), rot_y(th[0])));
    RI[1] = mul(mul(rot_z(psi[1]), mul(rot_x(phi[1]), rot_y(th[1]))), RI[0]);
    RI[2] = mul(rot_y(th[2]), RI[1]);
    RI[3] = mul(mul(rot_z(psi[3]), mul(rot_x(phi[3]), rot_y(th[3]))), RI[2]);
    RI[4] = mul(mul(rot_z(psi[4]), rot_y(th[4])), RI[3]);
    RI[5] = mul(mul(rot_z(psi[5]), rot_y(th[5])), RI[4]);
    RI[6] = mul(rot_y(th[6]), RI[2]);
    RI[7] = mul(rot_y(th[7]), RI[6]);
    RI[8] = mul(rot_y(th[8]), RI[2]);
    RI[9] = mul(rot_y(th[9]), RI[8]);
    RI[10] = mul(rot_y(th[10]), RI[3]);
    RI[11] = mul(rot_y(th[11]), RI[10]);
    RI[12] = mul(rot_y(th[12]), RI[3]);
    RI[13] = mul(rot_y(th[13]), RI[12]);
    for (int l = 0; l < NL; ++l) {
        double d[3];
        tmulv(&RI[seg[l].joint], seg[l].off, d);
        for (int i = 0; i < 3; ++i) P[l][i] = (seg[l].parent < 0 ? x[i] : P[seg[l].parent][i]) + d[i];
    }
    memset(Jfk, 0, sizeof(double) * NL * 3 * NA);
    for (int l = 0; l < NL; ++l)
        for (int i = 0; i < 3; ++i) Jfk[l][i][i] = 1.0;
    /* angle slots: (slot, joint, kind 0=theta 1=phi 2=psi) */
    static const int ang[22][3] = {
        {S_PHI0, 0, 1}, {S_PHI1, 1, 1}, {S_PHI3, 3, 1},
        {S_TH0 + 0, 0, 0}, {S_TH0 + 1, 1, 0}, {S_TH0 + 2, 2, 0}, {S_TH0 + 3, 3, 0}, {S_TH0 + 4, 4, 0},
        {S_TH0 + 5, 5, 0}, {S_TH0 + 6, 6, 0}, {S_TH0 + 7, 7, 0}, {S_TH0 + 8, 8, 0}, {S_TH0 + 9, 9, 0},
        {S_TH0 + 10, 10, 0}, {S_TH0 + 11, 11, 0}, {S_TH0 + 12, 12, 0}, {S_TH0 + 13, 13, 0},
        {S_PSI0, 0, 2}, {S_PSI1, 1, 2}, {S_PSI3, 3, 2}, {S_PSI4, 4, 2}, {S_PSI5, 5, 2}};
    for (int a = 0; a < 22; ++a) {
        const int slot = ang[a][0], k = ang[a][1], kind = ang[a][2];
        const int par = joint_parent[k];
        M3 I3 = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
        const M3* Rp = par >= 0 ? &RI[par] : &I3;
        double om[3];
        if (kind == 0) {            /* theta: R_parent_I e_y */
            const double ey[3] = {0, 1, 0};
            tmulv(Rp, ey, om);
        } else if (kind == 1) {     /* phi: R_parent_I Ry_a(theta) e_x ; Ry_a e_x = first row of rot_y */
            M3 ry = rot_y(th[k]);
            const double v[3] = {ry.m[0][0], ry.m[0][1], ry.m[0][2]};
            tmulv(Rp, v, om);
        } else {                    /* psi: R_k_I e_z */
            const double ez[3] = {0, 0, 1};
            tmulv(&RI[k], ez, om);
        }
        const int pv = joint_pivot[k];
        for (int l = 0; l < NL; ++l) {
            if (!is_ancestor(k, seg[l].joint)) continue;
            double d[3], c[3];
            for (int i = 0; i < 3; ++i) d[i] = P[l][i] - (pv < 0 ? x[i] : P[pv][i]);
            cross3(om, d, c);
            for (int i = 0; i < 3; ++i) Jfk[l][i][slot] = c[i];
        }
    }
}

static double step(double s, double x) { return 1.0 / (1.0 + exp(-(x - s))); }

/* rho, rho' and curvature weight psi = max(rho'/e, 1 - sigma_a) of the literal blend */
static void redesc(double e, double a, double b, double c, double* rho, double* d, double* psi) {
    const double sa = step(a, e), sb = step(b, e), sc = step(c, e);
    const double dsa = sa * (1 - sa), dsb = sb * (1 - sb), dsc = sc * (1 - sc);
    const double k3 = a * (c - b) / 2;
    const double p1 = e * e / 2, p2 = a * e - a * a / 2;
    const double u = (c - e) / (c - b);
    const double p3 = a * b - a * a / 2 + k3 * (1 - u * u), dp3 = a * u;
    const double p4 = a * b - a * a / 2 + k3;
    *rho = (1 - sa) * p1 + (sa - sb) * p2 + (sb - sc) * p3 + sc * p4;
    *d = -dsa * p1 + (1 - sa) * e + (dsa - dsb) * p2 + (sa - sb) * a + (dsb - dsc) * p3 + (sb - sc) * dp3 + dsc * p4;
    const double fl = 1 - sa;
    *psi = e > 0 ? fmax(*d / e, fl) : fl;
}

/* One frame. H is full 25x25 row-major (may be NULL). */
static void eval_frame(int C, const double* x, const double* meas, const double* w, const double* K,
                       const double* D, const double* R, const double* t, const double* abc, double* cost,
                       double* g, double* H) {
    double P[NL][3];
    static _Thread_local double Jfk[NL][3][NA];
    cheetah_fk(x, P, Jfk);
    double cst = 0;
    for (int i = 0; i < NA; ++i) g[i] = 0;
    if (H) memset(H, 0, sizeof(double) * NA * NA);
    for (int c = 0; c < C; ++c) {
        const double* Rc = R + 9 * c; const double* tc = t + 3 * c; const double* Dc = D + 4 * c;
        const double fx = K[9 * c + 0], fy = K[9 * c + 4], cx = K[9 * c + 2], cy = K[9 * c + 5];
        for (int l = 0; l < NL; ++l) {
            const double wt = w[c * NL + l];
            const double X = P[l][0], Y = P[l][1], Z = P[l][2];
            const double xc = Rc[0] * X + Rc[1] * Y + Rc[2] * Z + tc[0];
            const double yc = Rc[3] * X + Rc[4] * Y + Rc[5] * Z + tc[1];
            const double zc = Rc[6] * X + Rc[7] * Y + Rc[8] * Z + tc[2];
            const double a = xc / zc, b = yc / zc;
            const double r2 = a * a + b * b + 1e-12, r = sqrt(r2);
            const double th = atan(r), th2 = th * th;
            const double td = th * (1 + Dc[0] * th2 + Dc[1] * th2 * th2 + Dc[2] * th2 * th2 * th2 + Dc[3] * th2 * th2 * th2 * th2);
            const double dtd = 1 + 3 * Dc[0] * th2 + 5 * Dc[1] * th2 * th2 + 7 * Dc[2] * th2 * th2 * th2 + 9 * Dc[3] * th2 * th2 * th2 * th2;
            const double s = td / r;
            const double u = fx * a * s + cx, v = fy * b * s + cy;
            double res[2] = {u - meas[(c * NL + l) * 2 + 0], v - meas[(c * NL + l) * 2 + 1]};
            if (wt == 0.0) { res[0] = 0; res[1] = 0; }
            const double q = ((dtd * r / (1 + r * r) - td) / r2) / r;
            const double m00 = s + a * a * q, m01 = a * b * q, m11 = s + b * b * q;
            double Jc[2][3] = {{fx * m00 / zc, fx * m01 / zc, -fx * (m00 * a + m01 * b) / zc},
                               {fy * m01 / zc, fy * m11 / zc, -fy * (m01 * a + m11 * b) / zc}};
            for (int d2 = 0; d2 < 2; ++d2) {
                double Jw[3], Jrow[NA];
                for (int k = 0; k < 3; ++k) Jw[k] = Jc[d2][0] * Rc[k] + Jc[d2][1] * Rc[3 + k] + Jc[d2][2] * Rc[6 + k];
                for (int p = 0; p < NA; ++p) Jrow[p] = Jw[0] * Jfk[l][0][p] + Jw[1] * Jfk[l][1][p] + Jw[2] * Jfk[l][2][p];
                const double err = wt * res[d2];
                double rho, d, psi;
                redesc(fabs(err), abc[0], abc[1], abc[2], &rho, &d, &psi);
                cst += rho;
                if (wt == 0.0) continue;
                const double gam = (err < 0 ? -d : d) * wt, eta = psi * wt * wt;
                for (int p = 0; p < NA; ++p) g[p] += gam * Jrow[p];
                if (H)
                    for (int p = 0; p < NA; ++p) {
                        const double ep = eta * Jrow[p];
                        if (ep != 0.0)
                            for (int q2 = 0; q2 < NA; ++q2) H[p * NA + q2] += ep * Jrow[q2];
                    }
            }
        }
    }
    *cost = cst;
}

/* x [N][25], meas [N][C][20][2], w [N][C][20], K [C][9], D [C][4], R [C][9], t [C][3];
 * out: cost [N], g [N][25], H [N][25][25] (NULL to skip).  n_threads <= 0: all cores. */
void fte_oracle_eval(int n_frames, int n_cams, const double* x, const double* meas, const double* w,
                     const double* K, const double* D, const double* R, const double* t, const double* abc,
                     double* cost, double* g, double* H, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static)
#endif
    for (int n = 0; n < n_frames; ++n)
        eval_frame(n_cams, x + (size_t)n * NA, meas + (size_t)n * n_cams * NL * 2, w + (size_t)n * n_cams * NL, K, D,
                   R, t, abc, cost + n, g + (size_t)n * NA, H ? H + (size_t)n * NA * NA : 0);
}

int fte_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* marker positions only: pos [N][20][3] */
void fte_oracle_fk(int n_frames, const double* x, double* pos) {
    for (int n = 0; n < n_frames; ++n) {
        double P[NL][3];
        static _Thread_local double Jfk[NL][3][NA];
        cheetah_fk(x + (size_t)n * NA, P, Jfk);
        memcpy(pos + (size_t)n * NL * 3, P, sizeof(P));
    }
}
