// Pairwise extrinsic calibration of two fisheye cameras from checkerboard views, fp64.
//
// Replaces calibrate_pair_extrinsics_fisheye (/root/reference/src/calib/calib.py:125-134), i.e.
// cv2.fisheye.stereoCalibrate(..., flags=CALIB_FIX_INTRINSIC): minimise the reprojection error of the board corners in
// BOTH cameras over the relative pose (R_r, T) of camera 2 w.r.t. camera 1 and one board pose (R_v, t_v) per view,
//     Xc1 = R_v X + t_v,   Xc2 = R_r Xc1 + T,   e = [proj1(Xc1) - u1 ; proj2(Xc2) - u2].
// OpenCV (un-vendored dependency) parametrises rotations by Rodrigues vectors and runs plain Gauss-Newton from
// homography-based per-view poses; here rotations are updated multiplicatively, R <- R exp([d]x) (the Jacobian at d = 0
// is -R [X]x: no Rodrigues derivative), with Levenberg-Marquardt damping and the per-view 6 x 6 blocks eliminated
// (Schur complement onto the 6 relative-pose parameters) - the minimiser is the same, the path to it is not.
//
// Like skel_body.cuh the functions are __host__ __device__ loops over a context (tid, nthreads): csrc/stereo.cu runs
// them as CUDA kernels, tests/host_harness/stereo_host.cpp runs the same source on the CPU for the GPU-less test suite.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ST_HD __host__ __device__ __forceinline__
#else
#define ST_HD inline
#endif

namespace acino {

struct StereoCam {
    double fx, fy, cx, cy;
    double D[12];        // model 0 (fisheye): k1..k4;  model 1 (pinhole): k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (OpenCV order)
    int model;
};

constexpr int ST_POSE = 12;          // a pose record: R (9, row-major) + t (3)

ST_HD void st_mm(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// exp([w]x) (Rodrigues formula), row-major
ST_HD void st_exp(const double* w, double* R) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double a, b;                     // sin(th)/th, (1 - cos(th))/th^2
    if (th < 1e-6) {
        a = 1 - th2 / 6;
        b = 0.5 - th2 / 24;
    } else {
        a = sin(th) / th;
        b = (1 - cos(th)) / th2;
    }
    const double x = w[0], y = w[1], z = w[2];
    R[0] = 1 - b * (y * y + z * z); R[1] = -a * z + b * x * y;       R[2] = a * y + b * x * z;
    R[3] = a * z + b * x * y;       R[4] = 1 - b * (x * x + z * z);  R[5] = -a * x + b * y * z;
    R[6] = -a * y + b * x * z;      R[7] = a * x + b * y * z;        R[8] = 1 - b * (x * x + y * y);
}

// cv2.projectPoints model (plumb-bob / rational / thin-prism, calib.py:64-66) and its 2 x 3 Jacobian
ST_HD void st_project_pinhole(const StereoCam& cam, const double* Xc, double* uv, double* J) {
    const double* k = cam.D;
    const double iz = 1.0 / Xc[2], x = Xc[0] * iz, y = Xc[1] * iz;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double Nn = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6, Dn = 1 + k[5] * r2 + k[6] * r4 + k[7] * r6;
    const double cd = Nn / Dn;
    const double xd = x * cd + 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r4;
    const double yd = y * cd + k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r4;
    uv[0] = cam.fx * xd + cam.cx;
    uv[1] = cam.fy * yd + cam.cy;
    if (J) {
        const double dN = k[0] + 2 * k[1] * r2 + 3 * k[4] * r4, dD = k[5] + 2 * k[6] * r2 + 3 * k[7] * r4;
        const double dcd = (dN * Dn - Nn * dD) / (Dn * Dn);            // d cd / d r2
        const double sx = k[8] + 2 * k[9] * r2, sy = k[10] + 2 * k[11] * r2;
        const double xx = cd + 2 * x * x * dcd + 2 * k[2] * y + 6 * k[3] * x + 2 * x * sx;   // d xd / d x
        const double xy = 2 * x * y * dcd + 2 * k[2] * x + 2 * k[3] * y + 2 * y * sx;        // d xd / d y
        const double yx = 2 * x * y * dcd + 2 * k[2] * x + 2 * k[3] * y + 2 * x * sy;        // d yd / d x
        const double yy = cd + 2 * y * y * dcd + 6 * k[2] * y + 2 * k[3] * x + 2 * y * sy;   // d yd / d y
        const double fxi = cam.fx * iz, fyi = cam.fy * iz;
        J[0] = fxi * xx; J[1] = fxi * xy; J[2] = -fxi * (xx * x + xy * y);
        J[3] = fyi * yx; J[4] = fyi * yy; J[5] = -fyi * (yx * x + yy * y);
    }
}

// Kannala-Brandt projection of a camera-frame point and its 2 x 3 Jacobian (same model as acino_common.cuh)
ST_HD void st_project(const StereoCam& cam, const double* Xc, double* uv, double* J) {
    if (cam.model == 1) {
        st_project_pinhole(cam, Xc, uv, J);
        return;
    }
    const double iz = 1.0 / Xc[2], a = Xc[0] * iz, b = Xc[1] * iz;
    const double r2 = a * a + b * b + 1e-12, r = sqrt(r2), ir = 1.0 / r;
    const double th = atan(r), th2 = th * th;
    const double td = th * (1 + th2 * (cam.D[0] + th2 * (cam.D[1] + th2 * (cam.D[2] + th2 * cam.D[3]))));
    const double dtd = 1 + th2 * (3 * cam.D[0] + th2 * (5 * cam.D[1] + th2 * (7 * cam.D[2] + th2 * 9 * cam.D[3])));
    const double s = td * ir;
    uv[0] = cam.fx * a * s + cam.cx;
    uv[1] = cam.fy * b * s + cam.cy;
    if (J) {
        const double q = (dtd / (1 + r2) - s) * (ir * ir);
        const double m00 = s + a * a * q, m01 = a * b * q, m11 = s + b * b * q;
        const double fxi = cam.fx * iz, fyi = cam.fy * iz;
        J[0] = fxi * m00; J[1] = fxi * m01; J[2] = -fxi * (m00 * a + m01 * b);
        J[3] = fyi * m01; J[4] = fyi * m11; J[5] = -fyi * (m01 * a + m11 * b);
    }
}

// cv2.fisheye.undistortPoints (pixel -> normalised coordinates), Newton on theta_d(theta); returns false when it fails
ST_HD bool st_undistort(const StereoCam& cam, double u, double v, double* xy) {
    const double px = (u - cam.cx) / cam.fx, py = (v - cam.cy) / cam.fy;
    if (cam.model == 1) {            // fixed-point inversion of the pinhole distortion (initialisation only: 20 rounds)
        const double* k = cam.D;
        double x = px, y = py;
        for (int j = 0; j < 20; ++j) {
            const double r2 = x * x + y * y;
            const double icd = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icd < 0) return false;
            const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (px - dx) * icd;
            y = (py - dy) * icd;
        }
        xy[0] = x;
        xy[1] = y;
        return true;
    }
    double thd = sqrt(px * px + py * py);
    if (thd > 1.5707963267948966) thd = 1.5707963267948966;
    if (thd < 1e-8) { xy[0] = px; xy[1] = py; return true; }
    double th = thd;
    bool ok = false;
    for (int j = 0; j < 20; ++j) {
        const double t2 = th * th;
        const double f = th * (1 + t2 * (cam.D[0] + t2 * (cam.D[1] + t2 * (cam.D[2] + t2 * cam.D[3])))) - thd;
        const double df = 1 + t2 * (3 * cam.D[0] + t2 * (5 * cam.D[1] + t2 * (7 * cam.D[2] + t2 * 9 * cam.D[3])));
        const double fix = f / df;
        th -= fix;
        if (fabs(fix) < 1e-12) { ok = true; break; }
    }
    if (!ok || th <= 0) return false;
    const double sc = tan(th) / thd;
    xy[0] = px * sc;
    xy[1] = py * sc;
    return true;
}

// solve the SPD system A x = b (n <= 6) in place by Cholesky; returns false on a non-positive pivot.  A is overwritten.
ST_HD bool st_chol_solve(double* A, double* b, int n, int ld) {
    for (int j = 0; j < n; ++j) {
        double d = A[j * ld + j];
        for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
        if (!(d > 0)) return false;
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * ld + j];
            for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= A[i * ld + k] * b[k];
        b[i] = s / A[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= A[k * ld + i] * b[k];
        b[i] = s / A[i * ld + i];
    }
    return true;
}

// smallest eigenvector of a symmetric 9 x 9 matrix by cyclic Jacobi rotations (A destroyed); v [9]
ST_HD void st_min_eigvec9(double* A, double* v) {
    double Vm[81];
    for (int i = 0; i < 81; ++i) Vm[i] = (i % 10 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 9; ++p)
            for (int q = p + 1; q < 9; ++q) off += A[p * 9 + q] * A[p * 9 + q];
        if (off < 1e-300) break;
        for (int p = 0; p < 8; ++p)
            for (int q = p + 1; q < 9; ++q) {
                const double apq = A[p * 9 + q];
                if (fabs(apq) < 1e-300) continue;
                const double zeta = (A[q * 9 + q] - A[p * 9 + p]) / (2 * apq);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                const double c = 1 / sqrt(1 + tt * tt), s = c * tt;
                for (int k = 0; k < 9; ++k) {        // columns p, q
                    const double akp = A[k * 9 + p], akq = A[k * 9 + q];
                    A[k * 9 + p] = c * akp - s * akq;
                    A[k * 9 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 9; ++k) {        // rows p, q
                    const double apk = A[p * 9 + k], aqk = A[q * 9 + k];
                    A[p * 9 + k] = c * apk - s * aqk;
                    A[q * 9 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 9; ++k) {
                    const double vkp = Vm[k * 9 + p], vkq = Vm[k * 9 + q];
                    Vm[k * 9 + p] = c * vkp - s * vkq;
                    Vm[k * 9 + q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int j = 1; j < 9; ++j)
        if (A[j * 9 + j] < A[best * 9 + best]) best = j;
    for (int k = 0; k < 9; ++k) v[k] = Vm[k * 9 + best];
}

// Board pose of ONE view in ONE camera: planar homography (normalised DLT) -> [r1 r2 t] -> Gram-Schmidt -> a few
// damped Gauss-Newton steps on the reprojection error.  obj [M][3] (z = 0 plane), img [M][2] pixels -> pose [12].
// Returns the final sum of squared pixel errors (< 0: failed).
ST_HD double st_init_pose(const StereoCam& cam, int M, const double* obj, const double* img, double* pose) {
    // centroids / scales of both point sets (Hartley normalisation)
    double mx = 0, my = 0, mX = 0, mY = 0;
    int n = 0;
    for (int m = 0; m < M; ++m) {
        double xy[2];
        if (!st_undistort(cam, img[2 * m], img[2 * m + 1], xy)) continue;
        mx += xy[0]; my += xy[1]; mX += obj[3 * m]; mY += obj[3 * m + 1];
        ++n;
    }
    if (n < 4) return -1.0;
    mx /= n; my /= n; mX /= n; mY /= n;
    double sx = 0, sX = 0;
    for (int m = 0; m < M; ++m) {
        double xy[2];
        if (!st_undistort(cam, img[2 * m], img[2 * m + 1], xy)) continue;
        sx += sqrt((xy[0] - mx) * (xy[0] - mx) + (xy[1] - my) * (xy[1] - my));
        sX += sqrt((obj[3 * m] - mX) * (obj[3 * m] - mX) + (obj[3 * m + 1] - mY) * (obj[3 * m + 1] - mY));
    }
    sx = 1.4142135623730951 * n / sx;
    sX = 1.4142135623730951 * n / sX;
    double A[81];
    for (int i = 0; i < 81; ++i) A[i] = 0;
    for (int m = 0; m < M; ++m) {
        double xy[2];
        if (!st_undistort(cam, img[2 * m], img[2 * m + 1], xy)) continue;
        const double x = (xy[0] - mx) * sx, y = (xy[1] - my) * sx, X = (obj[3 * m] - mX) * sX, Y = (obj[3 * m + 1] - mY) * sX;
        const double r1[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, r2[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) A[i * 9 + j] += r1[i] * r1[j] + r2[i] * r2[j];
    }
    double hn[9];
    st_min_eigvec9(A, hn);
    // denormalise: H = Tx^-1 Hn TX, Tx = [sx 0 -sx mx; 0 sx -sx my; 0 0 1], TX likewise
    double H[9];
    {
        const double TX[9] = {sX, 0, -sX * mX, 0, sX, -sX * mY, 0, 0, 1};
        const double Ti[9] = {1 / sx, 0, mx, 0, 1 / sx, my, 0, 0, 1};
        double tmp[9];
        st_mm(hn, TX, tmp);
        st_mm(Ti, tmp, H);
    }
    double n1 = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
    double lam = 2.0 / (n1 + n2);
    if (H[8] * lam < 0) lam = -lam;                    // the board is in front of the camera: t_z > 0
    double r1[3] = {H[0] * lam, H[3] * lam, H[6] * lam}, r2[3] = {H[1] * lam, H[4] * lam, H[7] * lam};
    double t[3] = {H[2] * lam, H[5] * lam, H[8] * lam};
    n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    for (int i = 0; i < 3; ++i) r1[i] /= n1;
    const double d = r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2];
    for (int i = 0; i < 3; ++i) r2[i] -= d * r1[i];
    n2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    for (int i = 0; i < 3; ++i) r2[i] /= n2;
    const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    double R[9] = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]};
    // damped Gauss-Newton on the pixel error of this camera alone
    double cost = 0;
    double lm = 1e-3;
    for (int it = 0; it < 30; ++it) {
        double Hn[36], g[6];
        for (int i = 0; i < 36; ++i) Hn[i] = 0;
        for (int i = 0; i < 6; ++i) g[i] = 0;
        cost = 0;
        for (int m = 0; m < M; ++m) {
            const double* X = obj + 3 * m;
            const double RX[3] = {R[0] * X[0] + R[1] * X[1] + R[2] * X[2], R[3] * X[0] + R[4] * X[1] + R[5] * X[2],
                                  R[6] * X[0] + R[7] * X[1] + R[8] * X[2]};
            const double Xc[3] = {RX[0] + t[0], RX[1] + t[1], RX[2] + t[2]};
            double uv[2], Jp[6];
            st_project(cam, Xc, uv, Jp);
            // d Xc / d delta = -R [X]x
            double Jd[9];
            const double Xx[9] = {0, -X[2], X[1], X[2], 0, -X[0], -X[1], X[0], 0};
            st_mm(R, Xx, Jd);
            for (int dd = 0; dd < 2; ++dd) {
                const double e = uv[dd] - img[2 * m + dd];
                double J[6];
                for (int k = 0; k < 3; ++k) {
                    J[k] = -(Jp[3 * dd] * Jd[k] + Jp[3 * dd + 1] * Jd[3 + k] + Jp[3 * dd + 2] * Jd[6 + k]);
                    J[3 + k] = Jp[3 * dd + k];
                }
                cost += e * e;
                for (int i = 0; i < 6; ++i) {
                    g[i] += J[i] * e;
                    for (int j = 0; j <= i; ++j) Hn[i * 6 + j] += J[i] * J[j];
                }
            }
        }
        double Hd[36], step[6];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j <= i; ++j) Hd[i * 6 + j] = Hn[i * 6 + j];
            Hd[i * 6 + i] *= 1 + lm;
            step[i] = -g[i];
        }
        if (!st_chol_solve(Hd, step, 6, 6)) { lm *= 10; continue; }
        double Rn[9], E[9], tn[3] = {t[0] + step[3], t[1] + step[4], t[2] + step[5]};
        st_exp(step, E);
        st_mm(R, E, Rn);
        double cn = 0;
        for (int m = 0; m < M; ++m) {
            const double* X = obj + 3 * m;
            const double Xc[3] = {Rn[0] * X[0] + Rn[1] * X[1] + Rn[2] * X[2] + tn[0], Rn[3] * X[0] + Rn[4] * X[1] + Rn[5] * X[2] + tn[1],
                                  Rn[6] * X[0] + Rn[7] * X[1] + Rn[8] * X[2] + tn[2]};
            double uv[2];
            st_project(cam, Xc, uv, nullptr);
            cn += (uv[0] - img[2 * m]) * (uv[0] - img[2 * m]) + (uv[1] - img[2 * m + 1]) * (uv[1] - img[2 * m + 1]);
        }
        if (cn < cost) {
            const double rel = (cost - cn) / (cost > 1e-300 ? cost : 1e-300);
            for (int i = 0; i < 9; ++i) R[i] = Rn[i];
            for (int i = 0; i < 3; ++i) t[i] = tn[i];
            cost = cn;
            lm = lm > 1e-9 ? lm / 10 : lm;
            if (rel < 1e-12) break;
        } else {
            lm *= 10;
            if (lm > 1e8) break;
        }
    }
    for (int i = 0; i < 9; ++i) pose[i] = R[i];
    for (int i = 0; i < 3; ++i) pose[9 + i] = t[i];
    return cost;
}

// thread per (view, camera): pose_out [V][2][12], cost_out [V][2]
template <typename Ctx>
ST_HD void stereo_init_poses(const Ctx& ctx, const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj,
                             const double* img1, const double* img2, double* pose_out, double* cost_out) {
    for (int i = ctx.tid; i < 2 * V; i += ctx.nthreads) {
        const int v = i >> 1, c = i & 1;
        cost_out[i] = st_init_pose(c ? c2 : c1, M, obj, (c ? img2 : img1) + (size_t)v * M * 2, pose_out + (size_t)i * ST_POSE);
    }
}

// Per view: normal-equation blocks of the joint problem at (rel, poses), the per-view block eliminated.
//   out_cost [V]            sum of squared pixel errors of the view (both cameras)
//   out_S    [V][42]        Schur term of the view: 36 entries of W (V + lam diag V)^-1 W^T ... subtracted from U_v, i.e.
//                           S_v = U_v (1 + lam on its diagonal) - W_v Vd^-1 W_v^T   (row-major 6 x 6), then r_v (6) =
//                           -(gu_v - W_v Vd^-1 gv_v)
//   out_back [V][42]        Vd^-1 gv_v (6) and Vd^-1 W_v^T (6 x 6): d_v = -(Vd^-1 gv_v) - (Vd^-1 W_v^T) d_rel
// blocks == 0: cost only.  Parameter order: rotation increment (3), translation (3).
template <typename Ctx>
ST_HD void stereo_view_blocks(const Ctx& ctx, const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj,
                              const double* img1, const double* img2, const double* rel, const double* poses, double lam,
                              int blocks, double* out_cost, double* out_S, double* out_back, int* out_info) {
    for (int v = ctx.tid; v < V; v += ctx.nthreads) {
        const double* Rv = poses + (size_t)v * ST_POSE;
        const double* tv = Rv + 9;
        const double* Rr = rel;
        const double* Tr = rel + 9;
        double U[36], Wm[36], Vv[36], gu[6], gv[6], cost = 0;      // Wm[i][j]: rel param i x view param j
        for (int i = 0; i < 36; ++i) U[i] = Wm[i] = Vv[i] = 0;
        for (int i = 0; i < 6; ++i) gu[i] = gv[i] = 0;
        for (int m = 0; m < M; ++m) {
            const double* X = obj + 3 * m;
            const double RX[3] = {Rv[0] * X[0] + Rv[1] * X[1] + Rv[2] * X[2], Rv[3] * X[0] + Rv[4] * X[1] + Rv[5] * X[2],
                                  Rv[6] * X[0] + Rv[7] * X[1] + Rv[8] * X[2]};
            const double X1[3] = {RX[0] + tv[0], RX[1] + tv[1], RX[2] + tv[2]};
            const double X2[3] = {Rr[0] * X1[0] + Rr[1] * X1[1] + Rr[2] * X1[2] + Tr[0], Rr[3] * X1[0] + Rr[4] * X1[1] + Rr[5] * X1[2] + Tr[1],
                                  Rr[6] * X1[0] + Rr[7] * X1[1] + Rr[8] * X1[2] + Tr[2]};
            double uv1[2], uv2[2], P1[6], P2[6];
            st_project(c1, X1, uv1, blocks ? P1 : nullptr);
            st_project(c2, X2, uv2, blocks ? P2 : nullptr);
            const double e[4] = {uv1[0] - img1[((size_t)v * M + m) * 2], uv1[1] - img1[((size_t)v * M + m) * 2 + 1],
                                 uv2[0] - img2[((size_t)v * M + m) * 2], uv2[1] - img2[((size_t)v * M + m) * 2 + 1]};
            cost += e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
            if (!blocks) continue;
            // d X1 / d (dv, tv) = [-Rv [X]x | I];  d X2 / d (dv, tv) = Rr (that);  d X2 / d (dr, T) = [-Rr [X1]x | I]
            const double Xx[9] = {0, -X[2], X[1], X[2], 0, -X[0], -X[1], X[0], 0};
            const double X1x[9] = {0, -X1[2], X1[1], X1[2], 0, -X1[0], -X1[1], X1[0], 0};
            double A1[9], A2[9], B2[9];
            st_mm(Rv, Xx, A1);       // = Rv [X]x     (sign applied below)
            st_mm(Rr, A1, A2);       // = Rr Rv [X]x
            st_mm(Rr, X1x, B2);      // = Rr [X1]x
            for (int dd = 0; dd < 4; ++dd) {
                const bool cam2 = dd >= 2;
                const double* P = (cam2 ? P2 : P1) + 3 * (dd & 1);
                double Jv[6], Jr[6];
                for (int k = 0; k < 3; ++k) {
                    if (!cam2) {
                        Jv[k] = -(P[0] * A1[k] + P[1] * A1[3 + k] + P[2] * A1[6 + k]);
                        Jv[3 + k] = P[k];
                        Jr[k] = Jr[3 + k] = 0;
                    } else {
                        Jv[k] = -(P[0] * A2[k] + P[1] * A2[3 + k] + P[2] * A2[6 + k]);
                        Jv[3 + k] = P[0] * Rr[k] + P[1] * Rr[3 + k] + P[2] * Rr[6 + k];
                        Jr[k] = -(P[0] * B2[k] + P[1] * B2[3 + k] + P[2] * B2[6 + k]);
                        Jr[3 + k] = P[k];
                    }
                }
                for (int i = 0; i < 6; ++i) {
                    gv[i] += Jv[i] * e[dd];
                    for (int j = 0; j < 6; ++j) Vv[i * 6 + j] += Jv[i] * Jv[j];
                    if (cam2) {
                        gu[i] += Jr[i] * e[dd];
                        for (int j = 0; j < 6; ++j) {
                            U[i * 6 + j] += Jr[i] * Jr[j];
                            Wm[i * 6 + j] += Jr[i] * Jv[j];
                        }
                    }
                }
            }
        }
        out_cost[v] = cost;
        if (!blocks) continue;
        // Vd = V + lam diag V; solve Vd [y | Z] = [gv | W^T] column by column (one factorisation, 7 substitutions)
        double L[36];
        for (int i = 0; i < 36; ++i) L[i] = Vv[i];
        for (int i = 0; i < 6; ++i) L[i * 6 + i] *= 1 + lam;
        double y[6];
        for (int i = 0; i < 6; ++i) y[i] = gv[i];
        bool ok = st_chol_solve(L, y, 6, 6);            // L now holds the factor
        double Z[36];                                   // Z[j][i] = (Vd^-1 W^T)[j][i], W^T[j][i] = Wm[i][j]
        for (int i = 0; i < 6 && ok; ++i) {
            double col[6];
            for (int j = 0; j < 6; ++j) col[j] = Wm[i * 6 + j];
            // substitutions with the existing factor
            for (int a = 0; a < 6; ++a) {
                double s = col[a];
                for (int k = 0; k < a; ++k) s -= L[a * 6 + k] * col[k];
                col[a] = s / L[a * 6 + a];
            }
            for (int a = 5; a >= 0; --a) {
                double s = col[a];
                for (int k = a + 1; k < 6; ++k) s -= L[k * 6 + a] * col[k];
                col[a] = s / L[a * 6 + a];
            }
            for (int j = 0; j < 6; ++j) Z[j * 6 + i] = col[j];
        }
        if (!ok) {
            if (out_info) out_info[0] = v + 1;
            for (int i = 0; i < 42; ++i) out_S[(size_t)v * 42 + i] = out_back[(size_t)v * 42 + i] = 0;
            continue;
        }
        double* S = out_S + (size_t)v * 42;
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) {
                double s = U[i * 6 + j] * (i == j ? 1 + lam : 1.0);
                for (int k = 0; k < 6; ++k) s -= Wm[i * 6 + k] * Z[k * 6 + j];
                S[i * 6 + j] = s;
            }
            double r = gu[i];
            for (int k = 0; k < 6; ++k) r -= Wm[i * 6 + k] * y[k];
            S[36 + i] = -r;
        }
        double* Bk = out_back + (size_t)v * 42;
        for (int i = 0; i < 6; ++i) Bk[i] = y[i];
        for (int i = 0; i < 36; ++i) Bk[6 + i] = Z[i];
    }
}

// ONE thread: fixed-order sum of the per-view Schur terms, 6 x 6 solve -> d_rel [6]; info != 0: not positive definite
template <typename Ctx>
ST_HD void stereo_reduce_solve(const Ctx& ctx, int V, const double* S_all, double* d_rel, int* info) {
    if (ctx.tid != 0) return;
    double S[36], r[6];
    for (int i = 0; i < 36; ++i) S[i] = 0;
    for (int i = 0; i < 6; ++i) r[i] = 0;
    for (int v = 0; v < V; ++v) {
        for (int i = 0; i < 36; ++i) S[i] += S_all[(size_t)v * 42 + i];
        for (int i = 0; i < 6; ++i) r[i] += S_all[(size_t)v * 42 + 36 + i];
    }
    if (!st_chol_solve(S, r, 6, 6)) {
        info[0] = -1;
        for (int i = 0; i < 6; ++i) d_rel[i] = 0;
        return;
    }
    for (int i = 0; i < 6; ++i) d_rel[i] = r[i];
}

// trial point: rel' = (Rr exp(d_rel[0:3]), T + d_rel[3:6]); per view d_v = -y_v - Z_v d_rel, pose' likewise
template <typename Ctx>
ST_HD void stereo_update(const Ctx& ctx, int V, const double* rel, const double* poses, const double* back, const double* d_rel,
                         double* rel_t, double* poses_t) {
    for (int v = ctx.tid; v <= V; v += ctx.nthreads) {
        double d[6];
        const double* src;
        double* dst;
        if (v == V) {
            for (int i = 0; i < 6; ++i) d[i] = d_rel[i];
            src = rel;
            dst = rel_t;
        } else {
            const double* Bk = back + (size_t)v * 42;
            for (int i = 0; i < 6; ++i) {
                double s = -Bk[i];
                for (int k = 0; k < 6; ++k) s -= Bk[6 + i * 6 + k] * d_rel[k];
                d[i] = s;
            }
            src = poses + (size_t)v * ST_POSE;
            dst = poses_t + (size_t)v * ST_POSE;
        }
        double E[9], Rn[9];
        st_exp(d, E);
        st_mm(src, E, Rn);
        for (int i = 0; i < 9; ++i) dst[i] = Rn[i];
        for (int i = 0; i < 3; ++i) dst[9 + i] = src[9 + i] + d[3 + i];
    }
}

}  // namespace acino
