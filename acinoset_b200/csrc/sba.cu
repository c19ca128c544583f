// Sparse bundle adjustment of camera extrinsics + 3-D points (fp64).
//
// Replaces (reference, /root/reference/src/calib/calib.py):
//   params_to_points_extrinsics :345-352 (cv2.Rodrigues per camera)
//   cost_func_points_extrinsics :355-359 / cost_func_points_only :312-316 - the Python loop that
//       calls cv2.fisheye.projectPoints once per observation
//   SciPy's 2-point finite-difference Jacobian and TRF/LSMR step inside least_squares (:335,381)
// with: analytic per-observation blocks (Rodrigues derivative + fisheye Jacobian, SURVEY appendix
// B2), Cauchy IRLS weights rho'(z) = 1/(1+z), z = (f/C)^2 per residual coordinate (SciPy's loss),
// per-point 3x3 elimination (Schur complement) onto the 6C x 6C camera system, deterministic
// reductions (warp shuffles in fixed order -> per-CTA partials -> fixed-order final sum).
// Parameter layout = the reference's: [rvec_0..rvec_{C-1} | t_0..t_{C-1} | X_0..X_{n-1}].
#include "acino_common.cuh"
#include "stereo_body.cuh"

namespace acino {

struct SbaCam {           // per camera, refreshed from the parameter vector each evaluation
    double R[9];
    double dR[27];        // dR[i][j][k] = d R_ij / d rvec_k  at [ (i*3+j)*3 + k ]
    double t[3];
    // intrinsics + distortion: model 0 = Kannala-Brandt fisheye (k1..k4, cv2.fisheye.projectPoints, calib.py:132-136),
    // model 1 = OpenCV's standard model, up to 12 coefficients (cv2.projectPoints, calib.py:64-66)
    StereoCam in;
};

__device__ __forceinline__ void sba_set_intrinsics(SbaCam& cam, const int c, const int model, const int nd,
                                                   const double* __restrict__ K, const double* __restrict__ Dd) {
    cam.in.fx = K[9 * c]; cam.in.fy = K[9 * c + 4]; cam.in.cx = K[9 * c + 2]; cam.in.cy = K[9 * c + 5];
    for (int i = 0; i < 12; ++i) cam.in.D[i] = i < nd ? Dd[nd * c + i] : 0.0;
    cam.in.model = model;
}

// Rodrigues rotation and its derivative (closed form; generators at theta -> 0), thread per camera
__global__ void sba_cams_kernel(const int C, const int model, const int nd, const double* __restrict__ params,
                                const double* __restrict__ K, const double* __restrict__ Dd, SbaCam* __restrict__ cams) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double r[3] = {params[3 * c], params[3 * c + 1], params[3 * c + 2]};
    SbaCam cam;
    const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double Rm[9];
    if (th2 < 1e-30) {
        for (int i = 0; i < 9; ++i) Rm[i] = (i % 4 == 0) ? 1.0 : 0.0;
        for (int i = 0; i < 27; ++i) cam.dR[i] = 0.0;
        // generators: dR/dr_k = [e_k]x
        const int idx[3][2][2] = {{{2, 1}, {1, 2}}, {{0, 2}, {2, 0}}, {{1, 0}, {0, 1}}};
        for (int k = 0; k < 3; ++k) {
            cam.dR[(idx[k][0][0] * 3 + idx[k][0][1]) * 3 + k] = 1.0;
            cam.dR[(idx[k][1][0] * 3 + idx[k][1][1]) * 3 + k] = -1.0;
        }
    } else {
        const double th = sqrt(th2), ct = cos(th), st = sin(th);
        const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
        const double Kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Rm[i * 3 + j] = ct * (i == j) + (1 - ct) * k[i] * k[j] + st * Kx[i * 3 + j];
        // Gallego & Yezzi: dR/dr_k = (r_k [r]x + [r x (I - R) e_k]x) / |r|^2 * R
        const double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int kk = 0; kk < 3; ++kk) {
            double v[3];
            for (int i = 0; i < 3; ++i) v[i] = (i == kk ? 1.0 : 0.0) - Rm[i * 3 + kk];
            const double w[3] = {r[1] * v[2] - r[2] * v[1], r[2] * v[0] - r[0] * v[2], r[0] * v[1] - r[1] * v[0]};
            const double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
            double A[9];
            for (int i = 0; i < 9; ++i) A[i] = (r[kk] * rx[i] + wx[i]) / th2;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    cam.dR[(i * 3 + j) * 3 + kk] = A[i * 3] * Rm[j] + A[i * 3 + 1] * Rm[3 + j] + A[i * 3 + 2] * Rm[6 + j];
        }
    }
    for (int i = 0; i < 9; ++i) cam.R[i] = Rm[i];
    for (int i = 0; i < 3; ++i) cam.t[i] = params[3 * C + 3 * c + i];
    sba_set_intrinsics(cam, c, model, nd, K, Dd);
    cams[c] = cam;
}

// cameras given as fixed matrices (points-only mode): R, t straight from the scene
__global__ void sba_cams_fixed_kernel(const int C, const int model, const int nd, const double* __restrict__ R,
                                      const double* __restrict__ t, const double* __restrict__ K,
                                      const double* __restrict__ Dd, SbaCam* __restrict__ cams) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    SbaCam cam;
    for (int i = 0; i < 9; ++i) cam.R[i] = R[9 * c + i];
    for (int i = 0; i < 27; ++i) cam.dR[i] = 0.0;
    for (int i = 0; i < 3; ++i) cam.t[i] = t[3 * c + i];
    sba_set_intrinsics(cam, c, model, nd, K, Dd);
    cams[c] = cam;
}

// thread per observation: residual (2), Jc (2x6: d/d rvec, d/d t), Jp (2x3), Cauchy weights (2), cost.
// A thread's outputs are 23 doubles in five arrays at strides of 16 / 96 / 48 bytes: written directly, every store
// instruction of a warp touches 32 different sectors.  Full CTAs therefore stage their outputs in shared memory in
// the global layout (each array's slice of 128 observations is contiguous) and one thread issues five bulk async
// stores (TMA); the kernel's HBM traffic is then the algorithmic 176 B per observation.
constexpr int SBA_EVAL_THREADS = 128;
struct __align__(16) SbaEvalStage {
    double Jc[SBA_EVAL_THREADS * 12];
    double Jp[SBA_EVAL_THREADS * 6];
    double res[SBA_EVAL_THREADS * 2];
    double wgt[SBA_EVAL_THREADS * 2];
    double cost[SBA_EVAL_THREADS];
};

template <bool WANT_J>
__global__ void __launch_bounds__(SBA_EVAL_THREADS)
sba_eval_kernel(const int n_obs, const int bulk_ok, const SbaCam* __restrict__ cams, const double* __restrict__ pts,
                const float* __restrict__ uv, const int* __restrict__ cam_idx,
                const int* __restrict__ pt_idx, const double f_scale, double* __restrict__ res,
                double* __restrict__ Jc, double* __restrict__ Jp, double* __restrict__ wgt,
                double* __restrict__ cost) {
    __shared__ SbaEvalStage S;
    const int i0 = blockIdx.x * SBA_EVAL_THREADS;
    const int tid = threadIdx.x;
    const int i = i0 + tid;
    const bool staged = bulk_ok && i0 + SBA_EVAL_THREADS <= n_obs;      // uniform per CTA
    if (i < n_obs) {
        const SbaCam& cam = cams[cam_idx[i]];
        const double* X = pts + 3 * (size_t)pt_idx[i];
        const double x = X[0], y = X[1], z = X[2];
        const double xc = cam.R[0] * x + cam.R[1] * y + cam.R[2] * z + cam.t[0];
        const double yc = cam.R[3] * x + cam.R[4] * y + cam.R[5] * z + cam.t[1];
        const double zc = cam.R[6] * x + cam.R[7] * y + cam.R[8] * z + cam.t[2];
        ProjOut<double> pr;
        double ru, rv;
        if (cam.in.model == 1) {
            const double Xc[3] = {xc, yc, zc};
            double p2[2], J6[6];
            st_project_pinhole(cam.in, Xc, p2, WANT_J ? J6 : nullptr);
            ru = p2[0] - (double)uv[2 * i];
            rv = p2[1] - (double)uv[2 * i + 1];
            if (WANT_J) {
                for (int k = 0; k < 3; ++k) {
                    pr.ju[k] = J6[k];
                    pr.jv[k] = J6[3 + k];
                }
            }
        } else {
            fisheye_cam<double, WANT_J>(xc, yc, zc, cam.in.fx, cam.in.fy, cam.in.D[0], cam.in.D[1], cam.in.D[2], cam.in.D[3], pr);
            ru = pr.u + cam.in.cx - (double)uv[2 * i];
            rv = pr.v + cam.in.cy - (double)uv[2 * i + 1];
        }
        double* o_res = staged ? S.res + 2 * tid : res + 2 * (size_t)i;
        o_res[0] = ru;
        o_res[1] = rv;
        const double ic = 1.0 / f_scale;
        const double zu = ru * ic * ru * ic, zv = rv * ic * rv * ic;
        if (cost) (staged ? S.cost + tid : cost + i)[0] = 0.5 * f_scale * f_scale * (log1p(zu) + log1p(zv));
        if (WANT_J) {
            double* o_w = staged ? S.wgt + 2 * tid : wgt + 2 * (size_t)i;
            o_w[0] = 1.0 / (1.0 + zu);
            o_w[1] = 1.0 / (1.0 + zv);
            const double* J[2] = {pr.ju, pr.jv};
            if (Jc) {
                double* o_jc = staged ? S.Jc + 12 * tid : Jc + 12 * (size_t)i;
                // d Xc / d rvec_k = (dR/dr_k) X
                double M[3][3];
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 3; ++k)
                        M[a][k] = cam.dR[(a * 3 + 0) * 3 + k] * x + cam.dR[(a * 3 + 1) * 3 + k] * y + cam.dR[(a * 3 + 2) * 3 + k] * z;
                for (int d = 0; d < 2; ++d) {
                    for (int k = 0; k < 3; ++k) o_jc[d * 6 + k] = J[d][0] * M[0][k] + J[d][1] * M[1][k] + J[d][2] * M[2][k];
                    for (int k = 0; k < 3; ++k) o_jc[d * 6 + 3 + k] = J[d][k];
                }
            }
            double* o_jp = staged ? S.Jp + 6 * tid : Jp + 6 * (size_t)i;
            for (int d = 0; d < 2; ++d)
                for (int k = 0; k < 3; ++k) o_jp[d * 3 + k] = J[d][0] * cam.R[k] + J[d][1] * cam.R[3 + k] + J[d][2] * cam.R[6 + k];
        }
    }
    if (staged) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            auto st = [](void* dst, const void* src, unsigned bytes) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                             "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes)
                             : "memory");
            };
            st(res + 2 * (size_t)i0, S.res, SBA_EVAL_THREADS * 16);
            if (cost) st(cost + i0, S.cost, SBA_EVAL_THREADS * 8);
            if (WANT_J) {
                st(wgt + 2 * (size_t)i0, S.wgt, SBA_EVAL_THREADS * 16);
                if (Jc) st(Jc + 12 * (size_t)i0, S.Jc, SBA_EVAL_THREADS * 96);
                st(Jp + 6 * (size_t)i0, S.Jp, SBA_EVAL_THREADS * 48);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

// per-point pieces shared by the Schur and the back-substitution kernels
struct PointSys {
    double V[6];      // upper triangle of sum Jp^T w Jp: 00 01 02 11 12 22
    double gv[3];     // sum Jp^T w r
    double Vi[6];     // inverse of V + lam diag(V)
};

__device__ __forceinline__ void point_inverse(PointSys& s, const double lam) {     // Vi = (V + lam diag V)^-1
    const double a = s.V[0] * (1 + lam) + 1e-300, b = s.V[1], c = s.V[2], d = s.V[3] * (1 + lam) + 1e-300, e = s.V[4],
                 f = s.V[5] * (1 + lam) + 1e-300;
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    const double id = 1.0 / det;
    s.Vi[0] = c00 * id; s.Vi[1] = c01 * id; s.Vi[2] = c02 * id;
    s.Vi[3] = (a * f - c * c) * id; s.Vi[4] = (b * c - a * e) * id; s.Vi[5] = (a * d - b * b) * id;
}

__device__ __forceinline__ void point_system(const int o0, const int o1, const int* __restrict__ obs,
                                             const double* __restrict__ res, const double* __restrict__ Jp,
                                             const double* __restrict__ wgt, const double lam, PointSys& s) {
    for (int k = 0; k < 6; ++k) s.V[k] = 0.0;
    for (int k = 0; k < 3; ++k) s.gv[k] = 0.0;
    for (int o = o0; o < o1; ++o) {
        const int i = obs[o];
        for (int d = 0; d < 2; ++d) {
            const double w = wgt[2 * i + d], r = res[2 * i + d];
            const double* j = Jp + (size_t)i * 6 + d * 3;
            s.V[0] += w * j[0] * j[0]; s.V[1] += w * j[0] * j[1]; s.V[2] += w * j[0] * j[2];
            s.V[3] += w * j[1] * j[1]; s.V[4] += w * j[1] * j[2]; s.V[5] += w * j[2] * j[2];
            s.gv[0] += w * j[0] * r; s.gv[1] += w * j[1] * r; s.gv[2] += w * j[2] * r;
        }
    }
    point_inverse(s, lam);
}

// W_ap = Jc_a^T w Jp (6x3) for observation i
__device__ __forceinline__ void w_block(const int i, const double* __restrict__ Jc, const double* __restrict__ Jp,
                                        const double* __restrict__ wgt, double W[6][3]) {
    for (int r = 0; r < 6; ++r)
        for (int k = 0; k < 3; ++k) W[r][k] = 0.0;
    for (int d = 0; d < 2; ++d) {
        const double w = wgt[2 * i + d];
        const double* jc = Jc + (size_t)i * 12 + d * 6;
        const double* jp = Jp + (size_t)i * 6 + d * 3;
        for (int r = 0; r < 6; ++r)
            for (int k = 0; k < 3; ++k) W[r][k] += w * jc[r] * jp[k];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduced camera system.  One warp handles 32 points at a time (thread per point).  The point's observations are read
// once (ids and cameras kept in registers); the warp walks the cameras / camera pairs ANY of its points sees.  The 32
// per-point 6x6 contributions of a pair are summed across the warp by a transposition through shared memory: every lane
// writes its 36 values as one column of a [36][33] tile, then lane r adds up row r (and row r + 32) in lane order and
// adds the sum to ITS entries of the warp's private accumulator - fixed order, no shuffles, no serialised lane 0.
// (The first version reduced every value with a 5-step shuffle tree and let lane 0 accumulate: 1.7 ms per call at
// configs[3] against 0.1 ms for the evaluation.)  n = 6C; accumulator layout [S (n x n) | rhs (n) | diagU (n)].
constexpr int SCHUR_WARPS = 4;
constexpr int SCHUR_MAXOBS = 10;          // cameras per point <= cameras <= 10
constexpr int SCHUR_TILE_LD = 33;
__global__ void __launch_bounds__(SCHUR_WARPS * 32)
sba_schur_kernel(const int n_pts, const int C, const int* __restrict__ pt_ptr, const int* __restrict__ obs,
                 const int* __restrict__ cam_idx, const double* __restrict__ res, const double* __restrict__ Jc,
                 const double* __restrict__ Jp, const double* __restrict__ wgt, const double lam,
                 double* __restrict__ partial /*[grid][n*n + 2n]*/) {
    extern __shared__ double sacc[];     // [SCHUR_WARPS][n*n + 2n] accumulators, then [SCHUR_WARPS][36][33] tiles
    const int n = 6 * C, NR = n * n + 2 * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* acc = sacc + (size_t)warp * NR;
    double* tile = sacc + (size_t)SCHUR_WARPS * NR + (size_t)warp * 36 * SCHUR_TILE_LD;
    for (int i = lane; i < NR; i += 32) acc[i] = 0.0;
    __syncwarp();
    const int gw = blockIdx.x * SCHUR_WARPS + warp, nw = gridDim.x * SCHUR_WARPS;
    for (int base = gw * 32; base < n_pts; base += nw * 32) {
        const int p = base + lane;
        const bool live = p < n_pts;
        const int o0 = live ? pt_ptr[p] : 0;
        const int n_o = live ? min(pt_ptr[p + 1] - o0, SCHUR_MAXOBS) : 0;
        int oid[SCHUR_MAXOBS], ocam[SCHUR_MAXOBS];
        unsigned seen = 0;
#pragma unroll
        for (int k = 0; k < SCHUR_MAXOBS; ++k) {
            oid[k] = k < n_o ? obs[o0 + k] : -1;
            ocam[k] = k < n_o ? cam_idx[oid[k]] : -1;
            if (k < n_o) seen |= 1u << ocam[k];
        }
        // the point's 3x3 system (V + lam diag V)^-1 and gradient
        PointSys ps;
        for (int k = 0; k < 6; ++k) ps.V[k] = 0.0;
        for (int k = 0; k < 3; ++k) ps.gv[k] = 0.0;
#pragma unroll
        for (int k = 0; k < SCHUR_MAXOBS; ++k) {
            if (k >= n_o) continue;
            const int i = oid[k];
            for (int d = 0; d < 2; ++d) {
                const double w = wgt[2 * i + d], r = res[2 * i + d];
                const double* j = Jp + (size_t)i * 6 + d * 3;
                ps.V[0] += w * j[0] * j[0]; ps.V[1] += w * j[0] * j[1]; ps.V[2] += w * j[0] * j[2];
                ps.V[3] += w * j[1] * j[1]; ps.V[4] += w * j[1] * j[2]; ps.V[5] += w * j[2] * j[2];
                ps.gv[0] += w * j[0] * r; ps.gv[1] += w * j[1] * r; ps.gv[2] += w * j[2] * r;
            }
        }
        point_inverse(ps, lam);
        const unsigned any_seen = __reduce_or_sync(0xffffffffu, seen);
        for (int a = 0; a < C; ++a) {
            if (!((any_seen >> a) & 1u)) continue;
            int ia = -1;
#pragma unroll
            for (int k = 0; k < SCHUR_MAXOBS; ++k)
                if (ocam[k] == a) ia = oid[k];
            double Wa[6][3], Ya[6][3], ga[6], ja[2][6], wa[2] = {0.0, 0.0};
            for (int r = 0; r < 6; ++r) {
                ga[r] = 0.0;
                ja[0][r] = ja[1][r] = 0.0;
                for (int k = 0; k < 3; ++k) Wa[r][k] = Ya[r][k] = 0.0;
            }
            if (ia >= 0) {
                w_block(ia, Jc, Jp, wgt, Wa);
                for (int r = 0; r < 6; ++r) {
                    Ya[r][0] = Wa[r][0] * ps.Vi[0] + Wa[r][1] * ps.Vi[1] + Wa[r][2] * ps.Vi[2];
                    Ya[r][1] = Wa[r][0] * ps.Vi[1] + Wa[r][1] * ps.Vi[3] + Wa[r][2] * ps.Vi[4];
                    Ya[r][2] = Wa[r][0] * ps.Vi[2] + Wa[r][1] * ps.Vi[4] + Wa[r][2] * ps.Vi[5];
                }
                for (int d = 0; d < 2; ++d) {
                    wa[d] = wgt[2 * ia + d];
                    const double rr = res[2 * ia + d];
                    for (int r = 0; r < 6; ++r) {
                        ja[d][r] = Jc[(size_t)ia * 12 + d * 6 + r];
                        ga[r] += wa[d] * ja[d][r] * rr;
                    }
                }
            }
            // rhs_a = -g_a + Y_a gv (rows 0..5) and diag U_a (rows 6..11)
            for (int r = 0; r < 6; ++r) {
                tile[r * SCHUR_TILE_LD + lane] = -ga[r] + Ya[r][0] * ps.gv[0] + Ya[r][1] * ps.gv[1] + Ya[r][2] * ps.gv[2];
                tile[(6 + r) * SCHUR_TILE_LD + lane] = wa[0] * ja[0][r] * ja[0][r] + wa[1] * ja[1][r] * ja[1][r];
            }
            __syncwarp();
            if (lane < 12) {
                double sum = 0.0;
                for (int j = 0; j < 32; ++j) sum += tile[lane * SCHUR_TILE_LD + j];
                acc[n * n + (lane < 6 ? 6 * a + lane : n + 6 * a + lane - 6)] += sum;
            }
            __syncwarp();
            for (int b = a; b < C; ++b) {
                if (!((any_seen >> b) & 1u)) continue;
                int ib = -1;
                if (ia >= 0) {
#pragma unroll
                    for (int k = 0; k < SCHUR_MAXOBS; ++k)
                        if (ocam[k] == b) ib = oid[k];
                }
                if (!__any_sync(0xffffffffu, ib >= 0)) continue;
                double Wb[6][3];
                for (int r = 0; r < 6; ++r)
                    for (int k = 0; k < 3; ++k) Wb[r][k] = 0.0;
                if (ib >= 0) w_block(ib, Jc, Jp, wgt, Wb);
                for (int r = 0; r < 6; ++r)
                    for (int q = 0; q < 6; ++q) {
                        double v = -(Ya[r][0] * Wb[q][0] + Ya[r][1] * Wb[q][1] + Ya[r][2] * Wb[q][2]);
                        if (a == b) v += wa[0] * ja[0][r] * ja[0][q] + wa[1] * ja[1][r] * ja[1][q];
                        tile[(r * 6 + q) * SCHUR_TILE_LD + lane] = v;
                    }
                __syncwarp();
                for (int e = lane; e < 36; e += 32) {
                    double sum = 0.0;
                    for (int j = 0; j < 32; ++j) sum += tile[e * SCHUR_TILE_LD + j];
                    const int r = e / 6, q = e - 6 * r;
                    acc[(6 * a + r) * n + 6 * b + q] += sum;
                    if (a != b) acc[(6 * b + q) * n + 6 * a + r] += sum;
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NR; i += SCHUR_WARPS * 32) {
        double s = 0.0;
        for (int w = 0; w < SCHUR_WARPS; ++w) s += sacc[(size_t)w * NR + i];
        partial[(size_t)blockIdx.x * NR + i] = s;
    }
}

// fixed-order sum of the per-CTA partials (one warp per output entry: lane-strided partial sums, then a fixed shuffle
// tree); adds lam * diag(U) to the diagonal of S
__global__ void __launch_bounds__(128)
sba_schur_reduce_kernel(const int n_part, const int n, const double lam, const double* __restrict__ partial,
                        double* __restrict__ S, double* __restrict__ rhs) {
    const int NR = n * n + 2 * n;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n * n + n) return;
    const int r = i / n, c = i - r * n;
    const bool diag = i < n * n && r == c;
    double s = 0.0, du = 0.0;
    for (int b = lane; b < n_part; b += 32) {
        s += partial[(size_t)b * NR + i];
        if (diag) du += partial[(size_t)b * NR + n * n + n + r];
    }
    s = warp_sum(s);
    du = warp_sum(du);
    if (lane == 0) {
        if (i < n * n) S[i] = s + (diag ? lam * du : 0.0);
        else rhs[i - n * n] = s;
    }
}

// dense SPD solve S x = rhs for n <= 96 by Cholesky in one CTA (fp64); x overwrites rhs
__global__ void __launch_bounds__(128) sba_dense_solve_kernel(const int n, double* __restrict__ S, double* __restrict__ x,
                                                              int* __restrict__ info) {
    extern __shared__ double sm[];       // n x (n+1)
    const int ld = n + 1, tid = threadIdx.x;
    for (int i = tid; i < n * n; i += 128) sm[(i / n) * ld + (i % n)] = S[i];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = sm[k * ld + k];
            if (!(d > 0.0)) *info = k + 1;
            sm[k * ld + k] = sqrt(fmax(d, 1e-300));
        }
        __syncthreads();
        const double dk = sm[k * ld + k];
        for (int i = k + 1 + tid; i < n; i += 128) sm[i * ld + k] /= dk;
        __syncthreads();
        for (int t = tid; t < (n - k - 1) * (n - k - 1); t += 128) {
            const int i = k + 1 + t / (n - k - 1), j = k + 1 + t % (n - k - 1);
            if (j <= i) sm[i * ld + j] -= sm[i * ld + k] * sm[j * ld + k];
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int i = 0; i < n; ++i) {
            double s = x[i];
            for (int k = 0; k < i; ++k) s -= sm[i * ld + k] * x[k];
            x[i] = s / sm[i * ld + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = x[i];
            for (int k = i + 1; k < n; ++k) s -= sm[k * ld + i] * x[k];
            x[i] = s / sm[i * ld + i];
        }
    }
}

// thread per point: dp = -Vinv (gv + sum_c W_cp^T dc); writes the trial point and the model-reduction pieces
__global__ void sba_backsub_kernel(const int n_pts, const int C, const int* __restrict__ pt_ptr, const int* __restrict__ obs,
                                   const int* __restrict__ cam_idx, const double* __restrict__ res,
                                   const double* __restrict__ Jc, const double* __restrict__ Jp,
                                   const double* __restrict__ wgt, const double lam, const double* __restrict__ dc,
                                   const double* __restrict__ pts, double* __restrict__ pts_trial,
                                   double* __restrict__ dp_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const int o0 = pt_ptr[p], o1 = pt_ptr[p + 1];
    PointSys ps;
    point_system(o0, o1, obs, res, Jp, wgt, lam, ps);
    double v[3] = {ps.gv[0], ps.gv[1], ps.gv[2]};
    if (Jc && dc) {
        for (int o = o0; o < o1; ++o) {
            const int i = obs[o], a = cam_idx[i];
            double W[6][3];
            w_block(i, Jc, Jp, wgt, W);
            for (int r = 0; r < 6; ++r)
                for (int k = 0; k < 3; ++k) v[k] += W[r][k] * dc[6 * a + r];
        }
    }
    const double d0 = -(ps.Vi[0] * v[0] + ps.Vi[1] * v[1] + ps.Vi[2] * v[2]);
    const double d1 = -(ps.Vi[1] * v[0] + ps.Vi[3] * v[1] + ps.Vi[4] * v[2]);
    const double d2 = -(ps.Vi[2] * v[0] + ps.Vi[4] * v[1] + ps.Vi[5] * v[2]);
    pts_trial[3 * (size_t)p] = pts[3 * (size_t)p] + d0;
    pts_trial[3 * (size_t)p + 1] = pts[3 * (size_t)p + 1] + d1;
    pts_trial[3 * (size_t)p + 2] = pts[3 * (size_t)p + 2] + d2;
    if (dp_out) {
        dp_out[3 * (size_t)p] = d0; dp_out[3 * (size_t)p + 1] = d1; dp_out[3 * (size_t)p + 2] = d2;
    }
}

// model reduction of the step: thread per observation, pred_i = -(w r J d) - 1/2 w (J d)^2 summed over d
__global__ void sba_pred_kernel(const int n_obs, const int* __restrict__ cam_idx, const int* __restrict__ pt_idx,
                                const double* __restrict__ res, const double* __restrict__ Jc, const double* __restrict__ Jp,
                                const double* __restrict__ wgt, const double* __restrict__ dc, const double* __restrict__ dp,
                                double* __restrict__ pred) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_obs) return;
    const int a = cam_idx[i], p = pt_idx[i];
    double s = 0.0;
    for (int d = 0; d < 2; ++d) {
        double jd = 0.0;
        if (Jc && dc)
            for (int r = 0; r < 6; ++r) jd += Jc[(size_t)i * 12 + d * 6 + r] * dc[6 * a + r];
        for (int k = 0; k < 3; ++k) jd += Jp[(size_t)i * 6 + d * 3 + k] * dp[3 * (size_t)p + k];
        const double w = wgt[2 * i + d];
        s += -w * res[2 * i + d] * jd - 0.5 * w * jd * jd;
    }
    pred[i] = s;
}

// ---- launchers --------------------------------------------------------------------------------
static inline int nb(int n, int b) { return (n + b - 1) / b; }

cudaError_t launch_sba_cams(int C, int model, int n_dist, const double* params, const double* R, const double* t,
                            const double* K, const double* D, void* cams, cudaStream_t s) {
    if (params) sba_cams_kernel<<<1, 32, 0, s>>>(C, model, n_dist, params, K, D, (SbaCam*)cams);
    else sba_cams_fixed_kernel<<<1, 32, 0, s>>>(C, model, n_dist, R, t, K, D, (SbaCam*)cams);
    return cudaGetLastError();
}
size_t sba_cam_bytes() { return sizeof(SbaCam); }

cudaError_t launch_sba_eval(int n_obs, const void* cams, const double* pts, const float* uv, const int* cam_idx,
                            const int* pt_idx, double f_scale, double* res, double* Jc, double* Jp, double* wgt,
                            double* cost, cudaStream_t s) {
    if (n_obs <= 0) return cudaSuccess;
    const int bulk_ok = ((((uintptr_t)res | (uintptr_t)Jc | (uintptr_t)Jp | (uintptr_t)wgt | (uintptr_t)cost) & 15u) == 0) ? 1 : 0;
    if (Jp)
        sba_eval_kernel<true><<<nb(n_obs, SBA_EVAL_THREADS), SBA_EVAL_THREADS, 0, s>>>(n_obs, bulk_ok, (const SbaCam*)cams, pts, uv, cam_idx,
                                                                                     pt_idx, f_scale, res, Jc, Jp, wgt, cost);
    else
        sba_eval_kernel<false><<<nb(n_obs, SBA_EVAL_THREADS), SBA_EVAL_THREADS, 0, s>>>(n_obs, bulk_ok, (const SbaCam*)cams, pts, uv, cam_idx,
                                                                                      pt_idx, f_scale, res, Jc, Jp, wgt, cost);
    return cudaGetLastError();
}

int sba_schur_grid(int n_pts) {
    // one resident wave: 2 CTAs per SM by shared memory at 6 cameras (4 warps x (11 KB of accumulators + 9.5 KB of tile))
    int g = nb(n_pts, SCHUR_WARPS * 32);
    return g < 1 ? 1 : (g > 148 * 2 ? 148 * 2 : g);
}

cudaError_t launch_sba_schur(int n_pts, int C, const int* pt_ptr, const int* obs, const int* cam_idx, const double* res,
                             const double* Jc, const double* Jp, const double* wgt, double lam, double* partial, double* S,
                             double* rhs, cudaStream_t s) {
    const int n = 6 * C, NR = n * n + 2 * n, grid = sba_schur_grid(n_pts);
    const size_t smem = (size_t)SCHUR_WARPS * (NR + 36 * SCHUR_TILE_LD) * sizeof(double);
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(sba_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        set = true;
    }
    sba_schur_kernel<<<grid, SCHUR_WARPS * 32, smem, s>>>(n_pts, C, pt_ptr, obs, cam_idx, res, Jc, Jp, wgt, lam, partial);
    sba_schur_reduce_kernel<<<nb(n * n + n, 4), 128, 0, s>>>(grid, n, lam, partial, S, rhs);
    return cudaGetLastError();
}

cudaError_t launch_sba_dense_solve(int n, double* S, double* x, int* info, cudaStream_t s) {
    sba_dense_solve_kernel<<<1, 128, (size_t)n * (n + 1) * sizeof(double), s>>>(n, S, x, info);
    return cudaGetLastError();
}

cudaError_t launch_sba_backsub(int n_pts, int C, const int* pt_ptr, const int* obs, const int* cam_idx, const double* res,
                               const double* Jc, const double* Jp, const double* wgt, double lam, const double* dc,
                               const double* pts, double* pts_trial, double* dp, cudaStream_t s) {
    if (n_pts <= 0) return cudaSuccess;
    sba_backsub_kernel<<<nb(n_pts, 128), 128, 0, s>>>(n_pts, C, pt_ptr, obs, cam_idx, res, Jc, Jp, wgt, lam, dc, pts,
                                                      pts_trial, dp);
    return cudaGetLastError();
}

cudaError_t launch_sba_pred(int n_obs, const int* cam_idx, const int* pt_idx, const double* res, const double* Jc,
                            const double* Jp, const double* wgt, const double* dc, const double* dp, double* pred,
                            cudaStream_t s) {
    if (n_obs <= 0) return cudaSuccess;
    sba_pred_kernel<<<nb(n_obs, 128), 128, 0, s>>>(n_obs, cam_idx, pt_idx, res, Jc, Jp, wgt, dc, dp, pred);
    return cudaGetLastError();
}

}  // namespace acino
