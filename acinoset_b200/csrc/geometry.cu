// Camera-geometry kernels: fisheye projection of arbitrary 3-D points, OpenCV-compatible fisheye
// undistortion, two-view DLT triangulation and the adjacent-pair TRI driver.
//
// Replaces (reference, /root/reference/src/calib/calib.py):
//   project_points_fisheye            :132-136  (cv2.Rodrigues + cv2.fisheye.projectPoints)
//   triangulate_points_fisheye        :121-130  (cv2.fisheye.undistortPoints x2 + cv2.triangulatePoints)
//   get_pairwise_3d_points_from_df    :394-423  (adjacent pairs, inner merge, unweighted mean)
// OpenCV is an un-vendored dependency of the reference; the algorithms are restated from their
// published form and pinned to cv2 4.13.0 outputs in tests/golden/.
// All of this is fp64 by default: the reference returns float64 and the work is tiny; an fp32
// instantiation exists for the fused paths.  One thread per point; no reductions across threads,
// fixed pair order 0-1, 1-2, ... and fixed summation order => bit-reproducible.
#include "acino_common.cuh"

namespace acino {

template <typename T>
struct CamT {
    T R[9];
    T t[3];
    T fx, fy, cx, cy;
    T D[4];
};

template <typename T>
__device__ __forceinline__ CamT<T> load_cam(const CamD& c) {
    CamT<T> o;
#pragma unroll
    for (int i = 0; i < 9; ++i) o.R[i] = (T)c.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) o.t[i] = (T)c.t[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) o.D[i] = (T)c.D[i];
    o.fx = (T)c.fx; o.fy = (T)c.fy; o.cx = (T)c.cx; o.cy = (T)c.cy;
    return o;
}

// ---- projection -------------------------------------------------------------------------
template <typename T>
__global__ void project_points_kernel(const CamD camd, const int n, const T* __restrict__ X, T* __restrict__ uv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CamT<T> cam = load_cam<T>(camd);
    const T x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
    const T xc = cam.R[0] * x + cam.R[1] * y + cam.R[2] * z + cam.t[0];
    const T yc = cam.R[3] * x + cam.R[4] * y + cam.R[5] * z + cam.t[1];
    const T zc = cam.R[6] * x + cam.R[7] * y + cam.R[8] * z + cam.t[2];
    ProjOut<T> pr;
    fisheye_cam<T, false>(xc, yc, zc, cam.fx, cam.fy, cam.D[0], cam.D[1], cam.D[2], cam.D[3], pr);
    uv[2 * i] = pr.u + cam.cx;
    uv[2 * i + 1] = pr.v + cam.cy;
}

// ---- cv2.fisheye.undistortPoints (no R / P, default criteria: <= 10 Newton steps, eps 1e-8) ---
template <typename T>
__device__ __forceinline__ void undistort_point(const CamT<T>& cam, const T u, const T v, T& xo, T& yo) {
    const T pwx = (u - cam.cx) / cam.fx;
    const T pwy = (v - cam.cy) / cam.fy;
    T thd = sqrt(pwx * pwx + pwy * pwy);
    const T half_pi = T(1.5707963267948966);
    thd = fmin(fmax(-half_pi, thd), half_pi);
    const T eps = T(1e-8);
    bool converged = false;
    T th = thd;
    T scale = T(0);
    if (fabs(thd) > eps) {
        for (int j = 0; j < 10; ++j) {
            const T th2 = th * th, th4 = th2 * th2, th6 = th4 * th2, th8 = th6 * th2;
            const T k0 = cam.D[0] * th2, k1 = cam.D[1] * th4, k2 = cam.D[2] * th6, k3 = cam.D[3] * th8;
            const T fix = (th * (T(1) + k0 + k1 + k2 + k3) - thd) / (T(1) + T(3) * k0 + T(5) * k1 + T(7) * k2 + T(9) * k3);
            th = th - fix;
            if (fabs(fix) < eps) {
                converged = true;
                break;
            }
        }
        scale = tan(th) / thd;
    } else {
        converged = true;
    }
    const bool flipped = (thd < T(0) && th > T(0)) || (thd > T(0) && th < T(0));
    if (converged && !flipped) {
        xo = pwx * scale;
        yo = pwy * scale;
    } else {
        xo = T(-1000000.0);
        yo = T(-1000000.0);
    }
}

template <typename T>
__global__ void undistort_points_kernel(const CamD camd, const int n, const T* __restrict__ uv, T* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CamT<T> cam = load_cam<T>(camd);
    T x, y;
    undistort_point<T>(cam, uv[2 * i], uv[2 * i + 1], x, y);
    out[2 * i] = x;
    out[2 * i + 1] = y;
}

// ---- two-view DLT (cv2.triangulatePoints): smallest right singular vector of the 4x4 system,
//      by one-sided Jacobi (Hestenes) sweeps on the columns - no A^T A squaring ------------------
template <typename T>
__device__ __forceinline__ void dlt_pair(const CamT<T>& c1, const CamT<T>& c2, const T x1, const T y1, const T x2,
                                         const T y2, T X[3]) {
    T A[4][4], V[4][4];
    // rows: x P[2] - P[0], y P[2] - P[1] with P = [R | t], no row normalisation
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const T p10 = j < 3 ? c1.R[j] : c1.t[0], p11 = j < 3 ? c1.R[3 + j] : c1.t[1], p12 = j < 3 ? c1.R[6 + j] : c1.t[2];
        const T p20 = j < 3 ? c2.R[j] : c2.t[0], p21 = j < 3 ? c2.R[3 + j] : c2.t[1], p22 = j < 3 ? c2.R[6 + j] : c2.t[2];
        A[0][j] = x1 * p12 - p10;
        A[1][j] = y1 * p12 - p11;
        A[2][j] = x2 * p22 - p20;
        A[3][j] = y2 * p22 - p21;
#pragma unroll
        for (int i = 0; i < 4; ++i) V[i][j] = (i == j) ? T(1) : T(0);
    }
    const T tiny = sizeof(T) == 8 ? T(1e-300) : T(1e-30);
    for (int sweep = 0; sweep < 12; ++sweep) {
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                T al = T(0), be = T(0), ga = T(0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    al += A[i][p] * A[i][p];
                    be += A[i][q] * A[i][q];
                    ga += A[i][p] * A[i][q];
                }
                if (fabs(ga) > tiny) {
                    const T zeta = (be - al) / (T(2) * ga);
                    const T tt = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
                    const T c = T(1) / sqrt(T(1) + tt * tt);
                    const T s = c * tt;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const T ap = A[i][p], aq = A[i][q];
                        A[i][p] = c * ap - s * aq;
                        A[i][q] = s * ap + c * aq;
                        const T vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - s * vq;
                        V[i][q] = s * vp + c * vq;
                    }
                }
            }
    }
    int best = 0;
    T bn = T(0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T nn = T(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) nn += A[i][j] * A[i][j];
        if (j == 0 || nn < bn) {
            bn = nn;
            best = j;
        }
    }
    T v0 = V[0][0], v1 = V[1][0], v2 = V[2][0], v3 = V[3][0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (best == j) {
            v0 = V[0][j]; v1 = V[1][j]; v2 = V[2][j]; v3 = V[3][j];
        }
    X[0] = v0 / v3;
    X[1] = v1 / v3;
    X[2] = v2 / v3;
}

template <typename T>
__global__ void triangulate_points_kernel(const CamD cam1d, const CamD cam2d, const int n, const T* __restrict__ uv1,
                                          const T* __restrict__ uv2, T* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CamT<T> c1 = load_cam<T>(cam1d), c2 = load_cam<T>(cam2d);
    T x1, y1, x2, y2, X[3];
    undistort_point<T>(c1, uv1[2 * i], uv1[2 * i + 1], x1, y1);
    undistort_point<T>(c2, uv2[2 * i], uv2[2 * i + 1], x2, y2);
    dlt_pair<T>(c1, c2, x1, y1, x2, y2, X);
    out[3 * i] = X[0];
    out[3 * i + 1] = X[1];
    out[3 * i + 2] = X[2];
}

// ---- TRI driver: dense form of get_pairwise_3d_points_from_df ---------------------------------
struct CamTableD {
    CamD cam[ACINO_MAX_CAMS];
    int n_cams;
};

template <typename T>
__global__ void triangulate_pairwise_kernel(const __grid_constant__ CamTableD tab, const int n_frames, const int L,
                                            const T* __restrict__ uv, const unsigned char* __restrict__ valid,
                                            T* __restrict__ pos, int* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (frame, marker)
    if (i >= (long long)n_frames * L) return;
    const int n = (int)(i / L), l = (int)(i - (long long)n * L);
    const int C = tab.n_cams;
    T acc[3] = {T(0), T(0), T(0)};
    int cnt = 0;
    for (int c = 0; c + 1 < C; ++c) {
        const size_t ia = ((size_t)n * C + c) * L + l, ib = ((size_t)n * C + c + 1) * L + l;
        if (!valid[ia] || !valid[ib]) continue;
        const CamT<T> c1 = load_cam<T>(tab.cam[c]), c2 = load_cam<T>(tab.cam[c + 1]);
        T x1, y1, x2, y2, X[3];
        undistort_point<T>(c1, uv[2 * ia], uv[2 * ia + 1], x1, y1);
        undistort_point<T>(c2, uv[2 * ib], uv[2 * ib + 1], x2, y2);
        dlt_pair<T>(c1, c2, x1, y1, x2, y2, X);
        acc[0] += X[0];
        acc[1] += X[1];
        acc[2] += X[2];
        ++cnt;
    }
    const T nanv = sizeof(T) == 8 ? (T)__longlong_as_double(0x7ff8000000000000LL) : (T)__int_as_float(0x7fc00000);
    pos[3 * i] = cnt ? acc[0] / (T)cnt : nanv;
    pos[3 * i + 1] = cnt ? acc[1] / (T)cnt : nanv;
    pos[3 * i + 2] = cnt ? acc[2] / (T)cnt : nanv;
    if (count) count[i] = cnt;
}

// ---- pinhole twins (calib.py:52-66): cv2.projectPoints / cv2.undistortPoints with the plumb-bob, rational
//      (CALIB_RATIONAL_MODEL, calib.py:18) and thin-prism coefficients d = [k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4];
//      shorter vectors are zero-padded by the caller, the tilt terms (tauX, tauY) are rejected on the host ----
struct PinCamD {
    double R[9];
    double t[3];
    double fx, fy, cx, cy;
    double d[12];
};

__global__ void project_points_pinhole_kernel(const PinCamD cam, const int n, const double* __restrict__ X,
                                              double* __restrict__ uv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X0 = X[3 * i], X1 = X[3 * i + 1], X2 = X[3 * i + 2];
    const double xc = cam.R[0] * X0 + cam.R[1] * X1 + cam.R[2] * X2 + cam.t[0];
    const double yc = cam.R[3] * X0 + cam.R[4] * X1 + cam.R[5] * X2 + cam.t[1];
    double zc = cam.R[6] * X0 + cam.R[7] * X1 + cam.R[8] * X2 + cam.t[2];
    zc = zc != 0.0 ? 1.0 / zc : 1.0;                       // OpenCV: z = z ? 1/z : 1
    const double x = xc * zc, y = yc * zc;
    const double* k = cam.d;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    const double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
    const double icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
    const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
    const double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    uv[2 * i] = xd * cam.fx + cam.cx;
    uv[2 * i + 1] = yd * cam.fy + cam.cy;
}

// cv2.undistortPoints(pts, k, d[, P]) with its default criteria: exactly 5 fixed-point iterations, no
// convergence test; a negative inverse-distortion factor returns the plain normalised point
__device__ __forceinline__ void undistort_point_pinhole(const PinCamD& cam, const double u, const double v, double& xo,
                                                        double& yo) {
    const double* k = cam.d;
    double x = (u - cam.cx) / cam.fx, y = (v - cam.cy) / cam.fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {
            x = x0;
            y = y0;
            break;
        }
        const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    xo = x;
    yo = y;
}

// to_pixels: the `P=k` form of create_undistort_point_function (calib.py:25-30)
__global__ void undistort_points_pinhole_kernel(const PinCamD cam, const int n, const int to_pixels,
                                                const double* __restrict__ uv, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y;
    undistort_point_pinhole(cam, uv[2 * i], uv[2 * i + 1], x, y);
    out[2 * i] = to_pixels ? x * cam.fx + cam.cx : x;
    out[2 * i + 1] = to_pixels ? y * cam.fy + cam.cy : y;
}

__global__ void triangulate_points_pinhole_kernel(const PinCamD cam1, const PinCamD cam2, const int n,
                                                  const double* __restrict__ uv1, const double* __restrict__ uv2,
                                                  double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CamT<double> c1, c2;
#pragma unroll
    for (int j = 0; j < 9; ++j) { c1.R[j] = cam1.R[j]; c2.R[j] = cam2.R[j]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) { c1.t[j] = cam1.t[j]; c2.t[j] = cam2.t[j]; }
    double x1, y1, x2, y2, X[3];
    undistort_point_pinhole(cam1, uv1[2 * i], uv1[2 * i + 1], x1, y1);
    undistort_point_pinhole(cam2, uv2[2 * i], uv2[2 * i + 1], x2, y2);
    dlt_pair<double>(c1, c2, x1, y1, x2, y2, X);
    out[3 * i] = X[0];
    out[3 * i + 1] = X[1];
    out[3 * i + 2] = X[2];
}

// ---- launchers --------------------------------------------------------------------------------
static inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

cudaError_t launch_project_points_f64(const CamD& cam, int n, const double* X, double* uv, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    project_points_kernel<double><<<nblk(n, 128), 128, 0, s>>>(cam, n, X, uv);
    return cudaGetLastError();
}
cudaError_t launch_undistort_points_f64(const CamD& cam, int n, const double* uv, double* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    undistort_points_kernel<double><<<nblk(n, 128), 128, 0, s>>>(cam, n, uv, out);
    return cudaGetLastError();
}
cudaError_t launch_triangulate_points_f64(const CamD& c1, const CamD& c2, int n, const double* uv1, const double* uv2,
                                          double* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    triangulate_points_kernel<double><<<nblk(n, 64), 64, 0, s>>>(c1, c2, n, uv1, uv2, out);
    return cudaGetLastError();
}
cudaError_t launch_triangulate_pairwise_f64(const CamD* cams, int n_cams, int n_frames, int L, const double* uv,
                                            const unsigned char* valid, double* pos, int* count, cudaStream_t s) {
    if (n_frames <= 0 || L <= 0) return cudaSuccess;
    CamTableD tab;
    for (int c = 0; c < n_cams; ++c) tab.cam[c] = cams[c];
    tab.n_cams = n_cams;
    triangulate_pairwise_kernel<double><<<nblk((long long)n_frames * L, 64), 64, 0, s>>>(tab, n_frames, L, uv, valid, pos, count);
    return cudaGetLastError();
}

static PinCamD make_pincam(const double* K, const double* d, int nd, const double* R, const double* t) {
    PinCamD c;
    for (int i = 0; i < 9; ++i) c.R[i] = R ? R[i] : (i % 4 == 0 ? 1.0 : 0.0);
    for (int i = 0; i < 3; ++i) c.t[i] = t ? t[i] : 0.0;
    for (int i = 0; i < 12; ++i) c.d[i] = i < nd ? d[i] : 0.0;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    return c;
}
cudaError_t launch_project_points_pinhole(const double* K, const double* d, int nd, const double* R, const double* t, int n,
                                          const double* X, double* uv, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    project_points_pinhole_kernel<<<nblk(n, 128), 128, 0, s>>>(make_pincam(K, d, nd, R, t), n, X, uv);
    return cudaGetLastError();
}
cudaError_t launch_undistort_points_pinhole(const double* K, const double* d, int nd, int to_pixels, int n, const double* uv,
                                            double* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    undistort_points_pinhole_kernel<<<nblk(n, 128), 128, 0, s>>>(make_pincam(K, d, nd, nullptr, nullptr), n, to_pixels, uv, out);
    return cudaGetLastError();
}
cudaError_t launch_triangulate_points_pinhole(const double* K1, const double* d1, int nd1, const double* R1, const double* t1,
                                              const double* K2, const double* d2, int nd2, const double* R2, const double* t2,
                                              int n, const double* uv1, const double* uv2, double* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    triangulate_points_pinhole_kernel<<<nblk(n, 64), 64, 0, s>>>(make_pincam(K1, d1, nd1, R1, t1), make_pincam(K2, d2, nd2, R2, t2),
                                                                n, uv1, uv2, out);
    return cudaGetLastError();
}

}  // namespace acino

// ---- generic skeleton-pickle forward kinematics (reference src/build.py:32-95), fp64 ----------------
// Table driven: the host simulates the builder once (acinoset_b200/skeleton.py) and hands over, per
// link in order, (parent part, child part, whether the parent's LOCAL rotation is used transposed or
// not - the reference toggles `rot[child+"_i"]` every time a part appears as a child, build.py:78-79)
// and the rest-pose offset.  pose[child] = pose[parent] + M_parent * offset; roots sit at (x, y, z).
namespace acino {

constexpr int GFK_MAX_PARTS = 32;

__global__ void generic_fk_kernel(const int n_frames, const int n_parts, const int n_links, const int n_state,
                                  const int* __restrict__ dof_mask, const int* __restrict__ link_parent,
                                  const int* __restrict__ link_child, const int* __restrict__ link_flags,
                                  const double* __restrict__ link_tv, const double* __restrict__ x,
                                  double* __restrict__ pos) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_frames) return;
    const double* xs = x + (size_t)n * n_state;
    double pose[GFK_MAX_PARTS][3];
    for (int p = 0; p < n_parts; ++p) {
        pose[p][0] = xs[0];
        pose[p][1] = xs[1];
        pose[p][2] = xs[2];
    }
    for (int l = 0; l < n_links; ++l) {
        const int a = link_parent[l], b = link_child[l], fl = link_flags[l];
        // local passive rotation of part a: L = Rz(psi) Rx(phi) Ry(theta) (each factor only if its dof is set)
        double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        const int m = dof_mask[a];
        if (m & 2) {   // Ry(theta) (build.py:54-55)
            const double c = cos(xs[3 + n_parts + a]), s = sin(xs[3 + n_parts + a]);
            const double R[3][3] = {{c, 0, -s}, {0, 1, 0}, {s, 0, c}};
            double T[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * L[0][j] + R[i][1] * L[1][j] + R[i][2] * L[2][j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) L[i][j] = T[i][j];
        }
        if (m & 1) {   // Rx(phi) (build.py:56-57)
            const double c = cos(xs[3 + a]), s = sin(xs[3 + a]);
            const double R[3][3] = {{1, 0, 0}, {0, c, s}, {0, -s, c}};
            double T[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * L[0][j] + R[i][1] * L[1][j] + R[i][2] * L[2][j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) L[i][j] = T[i][j];
        }
        if (m & 4) {   // Rz(psi) (build.py:58-59)
            const double c = cos(xs[3 + 2 * n_parts + a]), s = sin(xs[3 + 2 * n_parts + a]);
            const double R[3][3] = {{c, s, 0}, {-s, c, 0}, {0, 0, 1}};
            double T[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * L[0][j] + R[i][1] * L[1][j] + R[i][2] * L[2][j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) L[i][j] = T[i][j];
        }
        const double* tv = link_tv + 3 * l;
        for (int i = 0; i < 3; ++i) {
            // flag 1: use L^T (the initial "_i" matrix), flag 0: use L (it has been transposed back)
            const double d = (fl & 1) ? (L[0][i] * tv[0] + L[1][i] * tv[1] + L[2][i] * tv[2])
                                      : (L[i][0] * tv[0] + L[i][1] * tv[1] + L[i][2] * tv[2]);
            pose[b][i] = pose[a][i] + d;
        }
    }
    for (int p = 0; p < n_parts; ++p)
        for (int i = 0; i < 3; ++i) pos[((size_t)n * n_parts + p) * 3 + i] = pose[p][i];
}

cudaError_t launch_generic_fk(int n_frames, int n_parts, int n_links, const int* dof_mask, const int* link_parent,
                              const int* link_child, const int* link_flags, const double* link_tv, const double* x,
                              double* pos, cudaStream_t s) {
    if (n_frames <= 0) return cudaSuccess;
    generic_fk_kernel<<<nblk(n_frames, 64), 64, 0, s>>>(n_frames, n_parts, n_links, 3 + 3 * n_parts, dof_mask, link_parent,
                                                        link_child, link_flags, link_tv, x, pos);
    return cudaGetLastError();
}

}  // namespace acino
