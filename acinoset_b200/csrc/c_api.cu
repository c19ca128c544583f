// C ABI of libacino_b200.so (see include/acino_b200.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "handle.cuh"

namespace acino {
cudaError_t launch_fte_eval(const SceneF& scene, int n_frames, const float* x, const float* meas,
                            const float* w, float* cost, float* g, float* H, cudaStream_t stream, int* sched);
const char* fte_eval_kernel_name(int n_frames);
cudaError_t launch_fk_project(const SceneF& scene, int n_frames, const float* x, float* pos, float* uv,
                              cudaStream_t stream);
cudaError_t launch_fte_jac(const SceneF& scene, int n_frames, const float* x, float* uv, float* J, cudaStream_t stream);
cudaError_t launch_project_points_f64(const CamD& cam, int n, const double* X, double* uv, cudaStream_t s);
cudaError_t launch_undistort_points_f64(const CamD& cam, int n, const double* uv, double* out, cudaStream_t s);
cudaError_t launch_triangulate_points_f64(const CamD& c1, const CamD& c2, int n, const double* uv1, const double* uv2,
                                          double* out, cudaStream_t s);
cudaError_t launch_triangulate_pairwise_f64(const CamD* cams, int n_cams, int n_frames, int L, const double* uv,
                                            const unsigned char* valid, double* pos, int* count, cudaStream_t s);
cudaError_t launch_project_points_pinhole(const double* K, const double* d, int nd, const double* R, const double* t, int n,
                                          const double* X, double* uv, cudaStream_t s);
cudaError_t launch_undistort_points_pinhole(const double* K, const double* d, int nd, int to_pixels, int n, const double* uv,
                                            double* out, cudaStream_t s);
cudaError_t launch_triangulate_points_pinhole(const double* K1, const double* d1, int nd1, const double* R1, const double* t1,
                                              const double* K2, const double* d2, int nd2, const double* R2, const double* t2,
                                              int n, const double* uv1, const double* uv2, double* out, cudaStream_t s);
cudaError_t launch_generic_fk(int n_frames, int n_parts, int n_links, const int* dof_mask, const int* link_parent,
                              const int* link_child, const int* link_flags, const double* link_tv, const double* x,
                              double* pos, cudaStream_t s);
size_t skel_eval_smem_bytes(int n_links, int n_out);
cudaError_t launch_skel_eval(const SkelDesc* d_skel, int n_links, int n_out, int P, int n_frames, const double* x,
                             const double* meas, const double* w, double* cost, double* g, double* H, cudaStream_t s);
cudaError_t launch_skel_prepare(int N, int P, int last_free, const double* x, const double* g, const double* sw,
                                const double* lo, const double* hi, double* gtot, unsigned char* fixed, double* cost_s,
                                cudaStream_t s);
cudaError_t launch_skel_assemble(int N, int P, const double* H, const double* gtot, const unsigned char* fixed,
                                 const double* sw, double lam, double* AB, double* rhs, cudaStream_t s);
cudaError_t launch_band_solve(long long n, int hb, double* AB, double* x, int* info, cudaStream_t s);
cudaError_t launch_skel_trial(int N, int P, int last_free, const double* x, const double* d, const double* lo,
                              const double* hi, double* xt, cudaStream_t s);
cudaError_t launch_skel_pred(int N, int P, const double* x, const double* xt, const double* gtot, const double* H,
                             const double* sw, double* pred, double* step, cudaStream_t s);
cudaError_t launch_stereo_init(const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj, const double* img1,
                               const double* img2, double* pose_out, double* cost_out, cudaStream_t s);
cudaError_t launch_stereo_blocks(const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj, const double* img1,
                                 const double* img2, const double* rel, const double* poses, double lam, int blocks,
                                 double* out_cost, double* out_S, double* out_back, int* out_info, cudaStream_t s);
cudaError_t launch_stereo_solve(int V, const double* S_all, double* d_rel, int* info, cudaStream_t s);
cudaError_t launch_stereo_update(int V, const double* rel, const double* poses, const double* back, const double* d_rel,
                                 double* rel_t, double* poses_t, cudaStream_t s);
cudaError_t launch_lm_prepare(int n_frames, long long frame0, long long ng, const double* x_ext, const float* g,
                              const double* sw, const double* lo, const double* hi, double* gtot, unsigned char* fixed,
                              double* cost_s, cudaStream_t s);
cudaError_t launch_lm_assemble(int n_frames, long long frame0, long long ng, int n_blocks, const float* H,
                               const double* gtot, const unsigned char* fixed, const double* sw, double lambda, double* D,
                               double* Lc, double* rhs, cudaStream_t s);
cudaError_t launch_lm_step(int n_frames, long long frame0, long long ng, const double* x_ext, const double* d_ext,
                           const double* gtot, const float* H, const double* sw, const double* lo, const double* hi,
                           double* xt_ext, float* xt32, double* pred, double* step, cudaStream_t s);
cudaError_t launch_lm_reduce(int n, const float* a0, const double* a1, const double* a2, const double* a3,
                             const double* m, double* out, double* ws, cudaStream_t s);
size_t lm_reduce_ws_bytes();
cudaError_t launch_bcr_factor(int n_elim, const int* elim, const double* D, const double* Lc, double* P, double* Q,
                              double* R, double* rhs, int* info, cudaStream_t s);
cudaError_t launch_bcr_update(int n_surv, const int* surv, double* D, double* Lc, const double* P, const double* Q,
                              double* rhs, cudaStream_t s);
cudaError_t launch_bcr_backsub(int n_elim, const int* elim, const double* R, const double* P, const double* Q,
                               const double* rhs, double* x, cudaStream_t s);
cudaError_t launch_sba_cams(int C, int model, int n_dist, const double* params, const double* R, const double* t,
                            const double* K, const double* D, void* cams, cudaStream_t s);
size_t sba_cam_bytes();
cudaError_t launch_sba_eval(int n_obs, const void* cams, const double* pts, const float* uv, const int* cam_idx,
                            const int* pt_idx, double f_scale, double* res, double* Jc, double* Jp, double* wgt,
                            double* cost, cudaStream_t s);
int sba_schur_grid(int n_pts);
cudaError_t launch_sba_schur(int n_pts, int C, const int* pt_ptr, const int* obs, const int* cam_idx, const double* res,
                             const double* Jc, const double* Jp, const double* wgt, double lam, double* partial, double* S,
                             double* rhs, cudaStream_t s);
cudaError_t launch_sba_dense_solve(int n, double* S, double* x, int* info, cudaStream_t s);
cudaError_t launch_sba_backsub(int n_pts, int C, const int* pt_ptr, const int* obs, const int* cam_idx, const double* res,
                               const double* Jc, const double* Jp, const double* wgt, double lam, const double* dc,
                               const double* pts, double* pts_trial, double* dp, cudaStream_t s);
cudaError_t launch_sba_pred(int n_obs, const int* cam_idx, const int* pt_idx, const double* res, const double* Jc,
                            const double* Jp, const double* wgt, const double* dc, const double* dp, double* pred,
                            cudaStream_t s);
}  // namespace acino

static void set_loss(LossF& L, double a, double b, double c) {
    L.a = (float)a; L.b = (float)b; L.c = (float)c;
    L.ea = (float)std::exp(a); L.eb = (float)std::exp(b); L.ec = (float)std::exp(c);
    L.p2c = (float)(-a * a / 2);
    L.k3 = (float)(a * (c - b) / 2);
    L.inv_cb = (float)(1.0 / (c - b));
    L.p3c = (float)(a * b - a * a / 2);
    L.p4 = (float)(a * b - a * a / 2 + a * (c - b) / 2);
    auto step = [](double s, double x) { return 1.0 / (1.0 + std::exp(-(x - s))); };
    const double sa = step(a, 0), sb = step(b, 0), sc = step(c, 0);
    const double u = c / (c - b);
    L.kappa = (float)(a / (2 * (c - b)));
    L.m2kappa = (float)(-a / (c - b));
    L.rho0 = (float)((sa - sb) * (-a * a / 2) + (sb - sc) * (a * b - a * a / 2 + (a * (c - b) / 2) * (1 - u * u)) +
                     sc * (a * b - a * a / 2 + a * (c - b) / 2));
}

static inline size_t pad64(size_t n) { return (n + 63) & ~(size_t)63; }

static int ensure_ws(acino_handle* h, size_t bytes) {
    if (bytes <= h->ws_bytes) return ACINO_OK;
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
    CK(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
    return ACINO_OK;
}

extern "C" {

int acino_version(void) { return 100; }

const char* acino_last_error(const acino_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int64_t acino_launch_count(const acino_handle* h) { return h ? h->launches : 0; }

int acino_create(acino_handle** out, int device) {
    if (!out) return fail(nullptr, ACINO_ERR_ARG, "acino_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, ACINO_ERR_CUDA, std::string("acino_create: no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, ACINO_ERR_ARG, "acino_create: bad device index");
    acino_handle* h = new acino_handle();
    h->device = device;
    memset(&h->scene, 0, sizeof(h->scene));
    set_loss(h->scene.loss, 3.0, 10.0, 20.0);   // all_optimizations.py:25-27
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->sched, 2 * acino_handle::kSchedSlots * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->sched, 0, 2 * acino_handle::kSchedSlots * sizeof(int));
    if (e != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "acino_create");
        delete h;
        return rc;
    }
    *out = h;
    return ACINO_OK;
}

int acino_destroy(acino_handle* h) {
    if (!h) return ACINO_OK;
    cudaSetDevice(h->device);
    if (h->ws) cudaFree(h->ws);
    if (h->red_ws) cudaFree(h->red_ws);
    if (h->sched) cudaFree(h->sched);
    if (h->d_skel) cudaFree(h->d_skel);
    if (h->st_buf) cudaFree(h->st_buf);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->pipe_ready) {
        cudaStreamDestroy(h->s_h2d);
        cudaStreamDestroy(h->s_d2h);
        for (int i = 0; i < acino_handle::kMaxChunks; ++i) {
            cudaEventDestroy(h->ev_in[i]);
            cudaEventDestroy(h->ev_done[i]);
        }
    }
    delete h;
    return ACINO_OK;
}

int acino_set_cameras(acino_handle* h, int n_cams, const double* K, const double* D, const double* R,
                      const double* t) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_set_cameras: NULL handle");
    if (n_cams < 1 || n_cams > ACINO_MAX_CAMS || !K || !D || !R || !t)
        return fail(h, ACINO_ERR_ARG, "acino_set_cameras: need 1..16 cameras and non-NULL K, D, R, t");
    for (int c = 0; c < n_cams; ++c) {
        CamD& cd = h->cam_d[c];
        CamF& cf = h->scene.cam[c];
        for (int i = 0; i < 9; ++i) { cd.R[i] = R[c * 9 + i]; cf.R[i] = (float)cd.R[i]; }
        for (int i = 0; i < 3; ++i) { cd.t[i] = t[c * 3 + i]; cf.t[i] = (float)cd.t[i]; }
        for (int i = 0; i < 4; ++i) { cd.D[i] = D[c * 4 + i]; cf.D[i] = (float)cd.D[i]; cf.D3[i] = (float)((2 * i + 3) * cd.D[i]); }
        cd.fx = K[c * 9 + 0]; cd.fy = K[c * 9 + 4]; cd.cx = K[c * 9 + 2]; cd.cy = K[c * 9 + 5];
        cf.fx = (float)cd.fx; cf.fy = (float)cd.fy; cf.cx = (float)cd.cx; cf.cy = (float)cd.cy;
    }
    // interleaved camera pairs for the packed-fp32 path; an odd last camera is paired with itself
    // (its second lane always runs with weight 0)
    for (int k = 0; k < (n_cams + 1) / 2; ++k) {
        const CamF& a = h->scene.cam[2 * k];
        const CamF& b = h->scene.cam[(2 * k + 1 < n_cams) ? 2 * k + 1 : 2 * k];
        CamPairF& p = h->scene.pair[k];
        for (int i = 0; i < 9; ++i) p.R[i] = make_float2(a.R[i], b.R[i]);
        for (int i = 0; i < 3; ++i) p.t[i] = make_float2(a.t[i], b.t[i]);
        for (int i = 0; i < 4; ++i) { p.D[i] = make_float2(a.D[i], b.D[i]); p.D3[i] = make_float2(a.D3[i], b.D3[i]); }
        p.fx = make_float2(a.fx, b.fx); p.fy = make_float2(a.fy, b.fy);
        p.cx = make_float2(a.cx, b.cx); p.cy = make_float2(a.cy, b.cy);
    }
    h->scene.n_cams = n_cams;
    h->have_cams = true;
    return ACINO_OK;
}

int acino_set_redescending(acino_handle* h, double a, double b, double c) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_set_redescending: NULL handle");
    if (!(a > 0 && b > a && c > b && c < 80.0))
        return fail(h, ACINO_ERR_ARG, "acino_set_redescending: need 0 < a < b < c < 80");
    set_loss(h->scene.loss, a, b, c);
    return ACINO_OK;
}

int acino_fte_eval_dev(acino_handle* h, int n_frames, const float* x, const float* meas, const float* w,
                       float* cost, float* g, float* H, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fte_eval_dev: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fte_eval_dev: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && (!x || !meas || !w)))
        return fail(h, ACINO_ERR_ARG, "acino_fte_eval_dev: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    if (((uintptr_t)meas & 7u) != 0) return fail(h, ACINO_ERR_ARG, "acino_fte_eval_dev: meas must be 8-byte aligned");
    CK(cudaSetDevice(h->device));
    CK(launch_fte_eval(h->scene, n_frames, x, meas, w, cost, g, H, (cudaStream_t)cuda_stream, h->next_sched()));
    h->launches += 1;
    return ACINO_OK;
}

const char* acino_fte_eval_kernel_name(int n_frames) { return fte_eval_kernel_name(n_frames); }

static int ensure_pipe(acino_handle* h) {
    if (h->pipe_ready) return ACINO_OK;
    CK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < acino_handle::kMaxChunks; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    h->pipe_ready = true;
    return ACINO_OK;
}

// Host-buffer entry point.  The batch is cut into chunks; chunk i+1 is copied in (stream s_h2d) while
// chunk i is evaluated (stream `stream`) and chunk i-1 is copied out (stream s_d2h) - PCIe is full
// duplex, so with pinned host buffers the call costs max(H2D, D2H) instead of H2D + kernel + D2H.
int acino_fte_eval(acino_handle* h, int n_frames, const float* x, const float* meas, const float* w,
                   float* cost, float* g, float* H) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fte_eval: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fte_eval: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && (!x || !meas || !w))) return fail(h, ACINO_ERR_ARG, "acino_fte_eval: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    int rc = ensure_pipe(h);
    if (rc) return rc;
    const size_t N = (size_t)n_frames, C = (size_t)h->scene.n_cams;
    const size_t nx = N * NA, nm = N * C * NL * 2, nw = N * C * NL, nc = N, ng = N * NA, nh = N * NU;
    // sub-buffers start on 64-float (256 B) boundaries: the kernel stages tiles with 16-byte bulk copies
    const size_t ax = pad64(nx), am = pad64(nm), aw = pad64(nw), ac = pad64(nc), ag = pad64(ng);
    const size_t total = (ax + am + aw + ac + ag + nh) * sizeof(float);
    rc = ensure_ws(h, total);
    if (rc) return rc;
    float* dx = (float*)h->ws;
    float* dm = dx + ax;
    float* dw = dm + am;
    float* dc = dw + aw;
    float* dg = dc + ac;
    float* dH = dg + ag;
    // chunk size: a multiple of 64 frames (keeps every chunk's tiles 16-byte aligned), <= kMaxChunks chunks
    size_t chunk = 16384;        // measured best of 4096..65536 on B200 (scripts/bench_e2e_chunks.sh)
#ifdef ACINO_EXPERIMENTS
    {
        const char* e = getenv("ACINO_E2E_CHUNK");      // A/B knob (frames per chunk, multiple of 64)
        const size_t c = e ? (size_t)atol(e) : 0;
        if (c >= 64 && c % 64 == 0) chunk = c;
    }
#endif
    while ((N + chunk - 1) / chunk > (size_t)acino_handle::kMaxChunks) chunk *= 2;
    const int n_chunks = (int)((N + chunk - 1) / chunk);
    for (int i = 0; i < n_chunks; ++i) {
        const size_t f0 = (size_t)i * chunk, nf = (f0 + chunk <= N) ? chunk : N - f0;
        CK(cudaMemcpyAsync(dx + f0 * NA, x + f0 * NA, nf * NA * sizeof(float), cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaMemcpyAsync(dm + f0 * C * NL * 2, meas + f0 * C * NL * 2, nf * C * NL * 2 * sizeof(float), cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaMemcpyAsync(dw + f0 * C * NL, w + f0 * C * NL, nf * C * NL * sizeof(float), cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaEventRecord(h->ev_in[i], h->s_h2d));
        CK(cudaStreamWaitEvent(h->stream, h->ev_in[i], 0));
        CK(launch_fte_eval(h->scene, (int)nf, dx + f0 * NA, dm + f0 * C * NL * 2, dw + f0 * C * NL, cost ? dc + f0 : nullptr,
                           g ? dg + f0 * NA : nullptr, H ? dH + f0 * NU : nullptr, h->stream, h->next_sched()));
        h->launches += 1;
        CK(cudaEventRecord(h->ev_done[i], h->stream));
        CK(cudaStreamWaitEvent(h->s_d2h, h->ev_done[i], 0));
        if (cost) CK(cudaMemcpyAsync(cost + f0, dc + f0, nf * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
        if (g) CK(cudaMemcpyAsync(g + f0 * NA, dg + f0 * NA, nf * NA * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
        if (H) CK(cudaMemcpyAsync(H + f0 * NU, dH + f0 * NU, nf * NU * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
    }
    CK(cudaStreamSynchronize(h->s_d2h));
    CK(cudaStreamSynchronize(h->stream));
    return ACINO_OK;
}

int acino_fk_project_dev(acino_handle* h, int n_frames, const float* x, float* pos, float* uv, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fk_project_dev: NULL handle");
    if (uv && !h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fk_project_dev: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && !x)) return fail(h, ACINO_ERR_ARG, "acino_fk_project_dev: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    CK(launch_fk_project(h->scene, n_frames, x, pos, uv, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_fk_project(acino_handle* h, int n_frames, const float* x, float* pos, float* uv) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fk_project: NULL handle");
    if (uv && !h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fk_project: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && !x)) return fail(h, ACINO_ERR_ARG, "acino_fk_project: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t N = (size_t)n_frames, C = (size_t)h->scene.n_cams;
    const size_t nx = N * NA, np = N * NL * 3, nu = N * C * NL * 2;
    int rc = ensure_ws(h, (pad64(nx) + pad64(np) + nu) * sizeof(float));
    if (rc) return rc;
    float* dx = (float*)h->ws;
    float* dp = dx + pad64(nx);
    float* du = dp + pad64(np);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dx, x, nx * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(launch_fk_project(h->scene, n_frames, dx, pos ? dp : nullptr, uv ? du : nullptr, s));
    h->launches += 1;
    if (pos) CK(cudaMemcpyAsync(pos, dp, np * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (uv) CK(cudaMemcpyAsync(uv, du, nu * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_fte_jac_dev(acino_handle* h, int n_frames, const float* x, float* uv, float* J, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fte_jac_dev: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fte_jac_dev: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && !x)) return fail(h, ACINO_ERR_ARG, "acino_fte_jac_dev: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    CK(launch_fte_jac(h->scene, n_frames, x, uv, J, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_fte_jac(acino_handle* h, int n_frames, const float* x, float* uv, float* J) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_fte_jac: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_fte_jac: cameras not set");
    if (n_frames < 0 || (n_frames > 0 && !x)) return fail(h, ACINO_ERR_ARG, "acino_fte_jac: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t N = (size_t)n_frames, C = (size_t)h->scene.n_cams;
    const size_t nx = N * NA, nu = N * C * NL * 2, nj = nu * NA;
    int rc = ensure_ws(h, (pad64(nx) + pad64(nu) + nj) * sizeof(float));
    if (rc) return rc;
    float* dx = (float*)h->ws;
    float* du = dx + pad64(nx);
    float* dj = du + pad64(nu);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dx, x, nx * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(launch_fte_jac(h->scene, n_frames, dx, uv ? du : nullptr, J ? dj : nullptr, s));
    h->launches += 1;
    if (uv) CK(cudaMemcpyAsync(uv, du, nu * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (J) CK(cudaMemcpyAsync(J, dj, nj * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

static CamD make_cam(const double* K, const double* D, const double* R, const double* t) {
    CamD c;
    for (int i = 0; i < 9; ++i) c.R[i] = R ? R[i] : (i % 4 == 0 ? 1.0 : 0.0);
    for (int i = 0; i < 3; ++i) c.t[i] = t ? t[i] : 0.0;
    for (int i = 0; i < 4; ++i) c.D[i] = D[i];
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    return c;
}

int acino_project_points(acino_handle* h, int n, const double* X, const double* K, const double* D, const double* R,
                         const double* t, double* uv) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_project_points: NULL handle");
    if (n < 0 || !K || !D || !R || !t || (n > 0 && (!X || !uv))) return fail(h, ACINO_ERR_ARG, "acino_project_points: bad arguments");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nx = pad64((size_t)n * 3), nu = (size_t)n * 2;
    int rc = ensure_ws(h, (nx + nu) * sizeof(double));
    if (rc) return rc;
    double* dX = (double*)h->ws;
    double* dU = dX + nx;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dX, X, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_project_points_f64(make_cam(K, D, R, t), n, dX, dU, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(uv, dU, nu * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_undistort_points(acino_handle* h, int n, const double* uv, const double* K, const double* D, double* xn) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_undistort_points: NULL handle");
    if (n < 0 || !K || !D || (n > 0 && (!uv || !xn))) return fail(h, ACINO_ERR_ARG, "acino_undistort_points: bad arguments");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nu = pad64((size_t)n * 2);
    int rc = ensure_ws(h, 2 * nu * sizeof(double));
    if (rc) return rc;
    double* dU = (double*)h->ws;
    double* dO = dU + nu;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dU, uv, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_undistort_points_f64(make_cam(K, D, nullptr, nullptr), n, dU, dO, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(xn, dO, (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_triangulate_points(acino_handle* h, int n, const double* uv1, const double* uv2, const double* K1,
                             const double* D1, const double* R1, const double* t1, const double* K2, const double* D2,
                             const double* R2, const double* t2, double* X) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_triangulate_points: NULL handle");
    if (n < 0 || !K1 || !D1 || !R1 || !t1 || !K2 || !D2 || !R2 || !t2 || (n > 0 && (!uv1 || !uv2 || !X)))
        return fail(h, ACINO_ERR_ARG, "acino_triangulate_points: bad arguments");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nu = pad64((size_t)n * 2);
    int rc = ensure_ws(h, (2 * nu + (size_t)n * 3) * sizeof(double));
    if (rc) return rc;
    double* d1 = (double*)h->ws;
    double* d2 = d1 + nu;
    double* dX = d2 + nu;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(d1, uv1, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d2, uv2, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_triangulate_points_f64(make_cam(K1, D1, R1, t1), make_cam(K2, D2, R2, t2), n, d1, d2, dX, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(X, dX, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

// ---- pinhole twins (calib.py:52-66)
static bool bad_dist(const double* d, int nd) {
    if (nd < 0 || nd > 14 || (nd > 0 && !d)) return true;
    for (int i = 12; i < nd; ++i)
        if (d[i] != 0.0) return true;      // tilted-sensor terms are not implemented
    return false;
}

int acino_project_points_pinhole(acino_handle* h, int n, const double* X, const double* K, const double* dist, int n_dist,
                                 const double* R, const double* t, double* uv) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_project_points_pinhole: NULL handle");
    if (n < 0 || !K || !R || !t || (n > 0 && (!X || !uv))) return fail(h, ACINO_ERR_ARG, "acino_project_points_pinhole: bad arguments");
    if (bad_dist(dist, n_dist)) return fail(h, ACINO_ERR_ARG, "acino_project_points_pinhole: 0..12 distortion coefficients (tilt terms must be 0)");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nx = pad64((size_t)n * 3), nu = (size_t)n * 2;
    int rc = ensure_ws(h, (nx + nu) * sizeof(double));
    if (rc) return rc;
    double* dX = (double*)h->ws;
    double* dU = dX + nx;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dX, X, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_project_points_pinhole(K, dist, n_dist < 12 ? n_dist : 12, R, t, n, dX, dU, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(uv, dU, nu * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_undistort_points_pinhole(acino_handle* h, int n, const double* uv, const double* K, const double* dist, int n_dist,
                                   int to_pixels, double* out) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_undistort_points_pinhole: NULL handle");
    if (n < 0 || !K || (n > 0 && (!uv || !out))) return fail(h, ACINO_ERR_ARG, "acino_undistort_points_pinhole: bad arguments");
    if (bad_dist(dist, n_dist)) return fail(h, ACINO_ERR_ARG, "acino_undistort_points_pinhole: 0..12 distortion coefficients (tilt terms must be 0)");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nu = pad64((size_t)n * 2);
    int rc = ensure_ws(h, 2 * nu * sizeof(double));
    if (rc) return rc;
    double* dU = (double*)h->ws;
    double* dO = dU + nu;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dU, uv, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_undistort_points_pinhole(K, dist, n_dist < 12 ? n_dist : 12, to_pixels, n, dU, dO, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(out, dO, (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_triangulate_points_pinhole(acino_handle* h, int n, const double* uv1, const double* uv2, const double* K1,
                                     const double* dist1, int n_dist1, const double* R1, const double* t1, const double* K2,
                                     const double* dist2, int n_dist2, const double* R2, const double* t2, double* X) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_triangulate_points_pinhole: NULL handle");
    if (n < 0 || !K1 || !R1 || !t1 || !K2 || !R2 || !t2 || (n > 0 && (!uv1 || !uv2 || !X)))
        return fail(h, ACINO_ERR_ARG, "acino_triangulate_points_pinhole: bad arguments");
    if (bad_dist(dist1, n_dist1) || bad_dist(dist2, n_dist2))
        return fail(h, ACINO_ERR_ARG, "acino_triangulate_points_pinhole: 0..12 distortion coefficients (tilt terms must be 0)");
    if (n == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nu = pad64((size_t)n * 2);
    int rc = ensure_ws(h, (2 * nu + (size_t)n * 3) * sizeof(double));
    if (rc) return rc;
    double* d1 = (double*)h->ws;
    double* d2 = d1 + nu;
    double* dX = d2 + nu;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(d1, uv1, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d2, uv2, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_triangulate_points_pinhole(K1, dist1, n_dist1 < 12 ? n_dist1 : 12, R1, t1, K2, dist2, n_dist2 < 12 ? n_dist2 : 12, R2, t2,
                                         n, d1, d2, dX, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(X, dX, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_triangulate_pairwise(acino_handle* h, int n_frames, int n_markers, const double* uv, const uint8_t* valid,
                               double* pos, int32_t* count) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_triangulate_pairwise: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_triangulate_pairwise: cameras not set");
    if (n_frames < 0 || n_markers < 0 || ((size_t)n_frames * n_markers > 0 && (!uv || !valid || !pos)))
        return fail(h, ACINO_ERR_ARG, "acino_triangulate_pairwise: bad arguments");
    const size_t N = (size_t)n_frames, L = (size_t)n_markers, C = (size_t)h->scene.n_cams;
    if (N * L == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t nu = pad64(N * C * L * 2), np = pad64(N * L * 3);
    const size_t nv = (N * C * L + 7) / 8 * 8, ncnt = N * L;
    int rc = ensure_ws(h, (nu + np) * sizeof(double) + ncnt * sizeof(int) + nv);
    if (rc) return rc;
    double* dU = (double*)h->ws;
    double* dP = dU + nu;
    int* dC = (int*)(dP + np);
    unsigned char* dV = (unsigned char*)(dC + ncnt);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dU, uv, N * C * L * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dV, valid, N * C * L, cudaMemcpyHostToDevice, s));
    CK(launch_triangulate_pairwise_f64(h->cam_d, (int)C, n_frames, n_markers, dU, dV, dP, dC, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(pos, dP, N * L * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (count) CK(cudaMemcpyAsync(count, dC, ncnt * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

int acino_generic_fk(acino_handle* h, int n_frames, int n_parts, int n_links, const int32_t* dof_mask,
                     const int32_t* link_parent, const int32_t* link_child, const int32_t* link_flags, const double* link_tv,
                     const double* x, double* pos) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_generic_fk: NULL handle");
    if (n_frames < 0 || n_parts < 1 || n_parts > 32 || n_links < 0 || !dof_mask || (n_links > 0 && (!link_parent || !link_child || !link_flags || !link_tv)) ||
        (n_frames > 0 && (!x || !pos)))
        return fail(h, ACINO_ERR_ARG, "acino_generic_fk: bad arguments (n_parts must be 1..32)");
    for (int l = 0; l < n_links; ++l)
        if (link_parent[l] < 0 || link_parent[l] >= n_parts || link_child[l] < 0 || link_child[l] >= n_parts)
            return fail(h, ACINO_ERR_ARG, "acino_generic_fk: link index out of range");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t ns = 3 + 3 * (size_t)n_parts, N = (size_t)n_frames;
    const size_t bx = N * ns * 8, bp = N * n_parts * 3 * 8, bi = (size_t)(n_parts + 3 * n_links) * 4, bt = (size_t)n_links * 3 * 8;
    const size_t ox = 0, op = (bx + 255) & ~(size_t)255, ot = op + ((bp + 255) & ~(size_t)255), oi = ot + ((bt + 255) & ~(size_t)255);
    int rc = ensure_ws(h, oi + bi + 256);
    if (rc) return rc;
    char* base = (char*)h->ws;
    double* dX = (double*)(base + ox);
    double* dP = (double*)(base + op);
    double* dT = (double*)(base + ot);
    int* dI = (int*)(base + oi);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dX, x, bx, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dI, dof_mask, (size_t)n_parts * 4, cudaMemcpyHostToDevice, s));
    if (n_links) {
        CK(cudaMemcpyAsync(dI + n_parts, link_parent, (size_t)n_links * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dI + n_parts + n_links, link_child, (size_t)n_links * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dI + n_parts + 2 * n_links, link_flags, (size_t)n_links * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dT, link_tv, bt, cudaMemcpyHostToDevice, s));
    }
    CK(launch_generic_fk(n_frames, n_parts, n_links, dI, dI + n_parts, dI + n_parts + n_links, dI + n_parts + 2 * n_links, dT, dX, dP, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(pos, dP, bp, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

#define DEV_ENTER(name)                                                         \
    if (!h) return fail(nullptr, ACINO_ERR_ARG, name ": NULL handle");          \
    CK(cudaSetDevice(h->device));                                               \
    cudaStream_t s = (cudaStream_t)cuda_stream

// ---- generic-skeleton FTE (build.py variant) ----------------------------------------------------------------
int acino_skel_set(acino_handle* h, int n_parts, int n_links, int n_out, const int32_t* dof_mask, const int32_t* link_parent,
                   const int32_t* link_flag, const double* link_tv, const uint64_t* path, int loss_kind, double a, double b,
                   double c, double delta) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_skel_set: NULL handle");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_skel_set: cameras not set (acino_set_cameras first)");
    if (n_parts < 1 || n_parts > SK_MAX_PARTS || n_links < 0 || n_links > SK_MAX_LINKS || n_out < 1 || n_out > SK_MAX_PARTS ||
        !dof_mask || !path || (n_links > 0 && (!link_parent || !link_flag || !link_tv)))
        return fail(h, ACINO_ERR_ARG, "acino_skel_set: need 1..32 parts / output rows, 0..40 links and non-NULL tables");
    if (loss_kind != 0 && loss_kind != 1) return fail(h, ACINO_ERR_ARG, "acino_skel_set: loss_kind is 0 (redescending) or 1 (abs)");
    if (loss_kind == 0 && !(a > 0 && b > a && c > b)) return fail(h, ACINO_ERR_ARG, "acino_skel_set: need 0 < a < b < c");
    if (loss_kind == 1 && !(delta > 0)) return fail(h, ACINO_ERR_ARG, "acino_skel_set: delta must be > 0");
    SkelDesc& S = h->skel;
    memset(&S, 0, sizeof(S));
    S.n_parts = n_parts; S.n_links = n_links; S.n_out = n_out; S.n_cams = h->scene.n_cams;
    for (int i = 0; i < n_parts; ++i) S.dof_mask[i] = dof_mask[i] & 7;
    for (int l = 0; l < n_links; ++l) {
        if (link_parent[l] < 0 || link_parent[l] >= n_parts) return fail(h, ACINO_ERR_ARG, "acino_skel_set: link parent out of range");
        S.link_parent[l] = link_parent[l];
        S.link_flag[l] = link_flag[l] & 1;
        for (int i = 0; i < 3; ++i) S.link_tv[l][i] = link_tv[3 * l + i];
    }
    for (int r = 0; r < n_out; ++r) {
        if (n_links < 64 && (path[r] >> n_links) != 0) return fail(h, ACINO_ERR_ARG, "acino_skel_set: path names a link that does not exist");
        S.path[r] = path[r];
    }
    int k = 0;
    for (int p = 0; p < n_parts; ++p) {
        S.part_ptr[p] = k;
        for (int l = 0; l < n_links; ++l)
            if (S.link_parent[l] == p) S.part_links[k++] = l;
    }
    for (int p = n_parts; p <= SK_MAX_PARTS; ++p) S.part_ptr[p] = k;
    S.loss_kind = loss_kind; S.la = a; S.lb = b; S.lc = c; S.delta = delta;
    for (int cI = 0; cI < S.n_cams; ++cI) {
        const CamD& cd = h->cam_d[cI];
        SkelCam& sc = S.cam[cI];
        for (int i = 0; i < 9; ++i) sc.R[i] = cd.R[i];
        for (int i = 0; i < 3; ++i) sc.t[i] = cd.t[i];
        for (int i = 0; i < 4; ++i) sc.D[i] = cd.D[i];
        sc.fx = cd.fx; sc.fy = cd.fy; sc.cx = cd.cx; sc.cy = cd.cy;
    }
    CK(cudaSetDevice(h->device));
    if (!h->d_skel) CK(cudaMalloc(&h->d_skel, sizeof(SkelDesc)));
    CK(cudaMemcpy(h->d_skel, &S, sizeof(SkelDesc), cudaMemcpyHostToDevice));
    h->have_skel = true;
    return ACINO_OK;
}

int acino_skel_eval_dev(acino_handle* h, int n_frames, const double* x, const double* meas, const double* w, double* cost,
                        double* g, double* H, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_skel_eval_dev: NULL handle");
    if (!h->have_skel) return fail(h, ACINO_ERR_STATE, "acino_skel_eval_dev: skeleton not set");
    if (n_frames < 0 || (n_frames > 0 && (!x || !meas || !w))) return fail(h, ACINO_ERR_ARG, "acino_skel_eval_dev: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    CK(launch_skel_eval(h->d_skel, h->skel.n_links, h->skel.n_out, 3 + 3 * h->skel.n_parts, n_frames, x, meas, w, cost, g, H,
                        (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_skel_eval(acino_handle* h, int n_frames, const double* x, const double* meas, const double* w, double* cost, double* g,
                    double* H) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_skel_eval: NULL handle");
    if (!h->have_skel) return fail(h, ACINO_ERR_STATE, "acino_skel_eval: skeleton not set");
    if (n_frames < 0 || (n_frames > 0 && (!x || !meas || !w))) return fail(h, ACINO_ERR_ARG, "acino_skel_eval: bad arguments");
    if (n_frames == 0) return ACINO_OK;
    CK(cudaSetDevice(h->device));
    const size_t N = (size_t)n_frames, P = 3 + 3 * (size_t)h->skel.n_parts, NU = P * (P + 1) / 2;
    const size_t CO = (size_t)h->skel.n_cams * h->skel.n_out;
    const size_t nx = pad64(N * P), nm = pad64(N * CO * 2), nw = pad64(N * CO), nc = pad64(N), ng = pad64(N * P);
    int rc = ensure_ws(h, (nx + nm + nw + nc + ng + N * NU) * sizeof(double));
    if (rc) return rc;
    double* dx = (double*)h->ws;
    double* dm = dx + nx;
    double* dw = dm + nm;
    double* dc = dw + nw;
    double* dg = dc + nc;
    double* dH = dg + ng;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dx, x, N * P * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dm, meas, N * CO * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dw, w, N * CO * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_skel_eval(h->d_skel, h->skel.n_links, h->skel.n_out, (int)P, n_frames, dx, dm, dw, cost ? dc : nullptr,
                        g ? dg : nullptr, H ? dH : nullptr, s));
    h->launches += 1;
    if (cost) CK(cudaMemcpyAsync(cost, dc, N * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (g) CK(cudaMemcpyAsync(g, dg, N * P * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (H) CK(cudaMemcpyAsync(H, dH, N * NU * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

#define SKEL_PRE(name)                                                                           \
    if (!h) return fail(nullptr, ACINO_ERR_ARG, name ": NULL handle");                           \
    if (!h->have_skel) return fail(h, ACINO_ERR_STATE, name ": skeleton not set");               \
    if (n_frames <= 0) return fail(h, ACINO_ERR_ARG, name ": n_frames must be positive");        \
    CK(cudaSetDevice(h->device));                                                                \
    const int P = 3 + 3 * h->skel.n_parts;

int acino_skel_prepare_dev(acino_handle* h, int n_frames, int last_free, const double* x, const double* g, const double* sw,
                           const double* lo, const double* hi, double* gtot, uint8_t* fixed, double* cost_s, void* cuda_stream) {
    SKEL_PRE("acino_skel_prepare_dev")
    if (!x || !g || !sw || !lo || !hi || !gtot || !fixed || !cost_s) return fail(h, ACINO_ERR_ARG, "acino_skel_prepare_dev: NULL pointer");
    CK(launch_skel_prepare(n_frames, P, last_free, x, g, sw, lo, hi, gtot, fixed, cost_s, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_skel_assemble_dev(acino_handle* h, int n_frames, const double* H, const double* gtot, const uint8_t* fixed,
                            const double* sw, double lambda, double* AB, double* rhs, void* cuda_stream) {
    SKEL_PRE("acino_skel_assemble_dev")
    if (!H || !gtot || !fixed || !sw || !AB || !rhs) return fail(h, ACINO_ERR_ARG, "acino_skel_assemble_dev: NULL pointer");
    CK(launch_skel_assemble(n_frames, P, H, gtot, fixed, sw, lambda, AB, rhs, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_band_solve_dev(acino_handle* h, int64_t n, int half_bandwidth, double* AB, double* x, int32_t* info, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_band_solve_dev: NULL handle");
    if (n <= 0 || half_bandwidth < 0 || !AB || !x || !info) return fail(h, ACINO_ERR_ARG, "acino_band_solve_dev: bad arguments");
    CK(cudaSetDevice(h->device));
    CK(launch_band_solve(n, half_bandwidth, AB, x, info, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_skel_trial_dev(acino_handle* h, int n_frames, int last_free, const double* x, const double* d, const double* lo,
                         const double* hi, double* xt, void* cuda_stream) {
    SKEL_PRE("acino_skel_trial_dev")
    if (!x || !d || !lo || !hi || !xt) return fail(h, ACINO_ERR_ARG, "acino_skel_trial_dev: NULL pointer");
    CK(launch_skel_trial(n_frames, P, last_free, x, d, lo, hi, xt, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

int acino_skel_pred_dev(acino_handle* h, int n_frames, const double* x, const double* xt, const double* gtot, const double* H,
                        const double* sw, double* pred, double* step, void* cuda_stream) {
    SKEL_PRE("acino_skel_pred_dev")
    if (!x || !xt || !gtot || !H || !sw || !pred || !step) return fail(h, ACINO_ERR_ARG, "acino_skel_pred_dev: NULL pointer");
    CK(launch_skel_pred(n_frames, P, x, xt, gtot, H, sw, pred, step, (cudaStream_t)cuda_stream));
    h->launches += 1;
    return ACINO_OK;
}

// ---- pairwise fisheye extrinsic calibration (calib.py:125-134), host pointers --------------------
static StereoCam make_stcam(const double* K, const double* D, int nd, int model) {
    StereoCam c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    for (int i = 0; i < 12; ++i) c.D[i] = i < nd ? D[i] : 0.0;
    c.model = model;
    return c;
}

static int stereo_set_impl(acino_handle* h, int n_views, int n_points, const double* obj, const double* img1, const double* img2,
                           const double* K1, const double* D1, int nd1, const double* K2, const double* D2, int nd2, int model);

int acino_stereo_set(acino_handle* h, int n_views, int n_points, const double* obj, const double* img1, const double* img2,
                     const double* K1, const double* D1, const double* K2, const double* D2) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_stereo_set: NULL handle");
    if (!D1 || !D2) return fail(h, ACINO_ERR_ARG, "acino_stereo_set: NULL distortion");
    return stereo_set_impl(h, n_views, n_points, obj, img1, img2, K1, D1, 4, K2, D2, 4, 0);
}

int acino_stereo_set_pinhole(acino_handle* h, int n_views, int n_points, const double* obj, const double* img1, const double* img2,
                             const double* K1, const double* dist1, int n_dist1, const double* K2, const double* dist2, int n_dist2) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_stereo_set_pinhole: NULL handle");
    if (bad_dist(dist1, n_dist1) || bad_dist(dist2, n_dist2))
        return fail(h, ACINO_ERR_ARG, "acino_stereo_set_pinhole: 0..12 distortion coefficients (tilt terms must be 0)");
    return stereo_set_impl(h, n_views, n_points, obj, img1, img2, K1, dist1, n_dist1 < 12 ? n_dist1 : 12, K2, dist2,
                           n_dist2 < 12 ? n_dist2 : 12, 1);
}

static int stereo_set_impl(acino_handle* h, int n_views, int n_points, const double* obj, const double* img1, const double* img2,
                           const double* K1, const double* D1, int nd1, const double* K2, const double* D2, int nd2, int model) {
    if (n_views < 1 || n_points < 4 || !obj || !img1 || !img2 || !K1 || !K2)
        return fail(h, ACINO_ERR_ARG, "acino_stereo_set: need >= 1 view, >= 4 points per view and non-NULL arrays");
    CK(cudaSetDevice(h->device));
    const size_t V = (size_t)n_views, M = (size_t)n_points;
    const size_t n_obj = pad64(M * 3), n_img = pad64(V * M * 2), n_pose = pad64(V * 12 * 2), n_view = pad64(V * 42);
    const size_t total = n_obj + 2 * n_img + 64 + 64 + 2 * n_pose + pad64(2 * V) + 2 * n_view + 64 + 64;
    if (h->st_buf) cudaFree(h->st_buf);
    h->st_buf = nullptr;
    CK(cudaMalloc((void**)&h->st_buf, total * sizeof(double)));
    double* p = h->st_buf;
    h->st_obj = p; p += n_obj;
    h->st_img1 = p; p += n_img;
    h->st_img2 = p; p += n_img;
    h->st_rel = p; p += 64;
    h->st_rel_t = p; p += 64;
    h->st_poses = p; p += n_pose;
    h->st_poses_t = p; p += n_pose;
    h->st_cost = p; p += pad64(2 * V);
    h->st_S = p; p += n_view;
    h->st_back = p; p += n_view;
    h->st_drel = p; p += 64;
    h->st_info = (int*)p;
    CK(cudaMemcpy(h->st_obj, obj, M * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->st_img1, img1, V * M * 2 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->st_img2, img2, V * M * 2 * sizeof(double), cudaMemcpyHostToDevice));
    h->st_c1 = make_stcam(K1, D1, nd1, model);
    h->st_c2 = make_stcam(K2, D2, nd2, model);
    h->st_V = n_views;
    h->st_M = n_points;
    return ACINO_OK;
}

int acino_stereo_init(acino_handle* h, double* poses, double* cost) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_stereo_init: NULL handle");
    if (!h->st_V) return fail(h, ACINO_ERR_STATE, "acino_stereo_init: no problem set (acino_stereo_set first)");
    if (!poses || !cost) return fail(h, ACINO_ERR_ARG, "acino_stereo_init: NULL output");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t V = (size_t)h->st_V;
    CK(launch_stereo_init(h->st_c1, h->st_c2, h->st_V, h->st_M, h->st_obj, h->st_img1, h->st_img2, h->st_poses, h->st_cost, s));
    h->launches += 1;
    CK(cudaMemcpyAsync(poses, h->st_poses, V * 2 * 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(cost, h->st_cost, V * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ACINO_OK;
}

// lambda < 0: cost of (rel, poses) only.  Otherwise one damped Gauss-Newton step from (rel, poses): trial point
// (rel_t, poses_t) and its cost.  cost = sum of squared pixel errors over both cameras, summed in view order.
int acino_stereo_step(acino_handle* h, const double* rel, const double* poses, double lambda, double* rel_t, double* poses_t,
                      double* cost, int32_t* info) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_stereo_step: NULL handle");
    if (!h->st_V) return fail(h, ACINO_ERR_STATE, "acino_stereo_step: no problem set (acino_stereo_set first)");
    if (!rel || !poses || !cost || !info || (lambda >= 0 && (!rel_t || !poses_t)))
        return fail(h, ACINO_ERR_ARG, "acino_stereo_step: NULL pointer");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int V = h->st_V, M = h->st_M;
    CK(cudaMemcpyAsync(h->st_rel, rel, 12 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->st_poses, poses, (size_t)V * 12 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->st_info, 0, 2 * sizeof(int), s));
    const double* rel_eval = h->st_rel;
    const double* poses_eval = h->st_poses;
    if (lambda >= 0) {
        CK(launch_stereo_blocks(h->st_c1, h->st_c2, V, M, h->st_obj, h->st_img1, h->st_img2, h->st_rel, h->st_poses, lambda, 1,
                                h->st_cost, h->st_S, h->st_back, h->st_info, s));
        CK(launch_stereo_solve(V, h->st_S, h->st_drel, h->st_info + 1, s));
        CK(launch_stereo_update(V, h->st_rel, h->st_poses, h->st_back, h->st_drel, h->st_rel_t, h->st_poses_t, s));
        h->launches += 3;
        rel_eval = h->st_rel_t;
        poses_eval = h->st_poses_t;
    }
    CK(launch_stereo_blocks(h->st_c1, h->st_c2, V, M, h->st_obj, h->st_img1, h->st_img2, rel_eval, poses_eval, 0.0, 0, h->st_cost,
                            nullptr, nullptr, nullptr, s));
    h->launches += 1;
    std::vector<double> cv((size_t)V);
    int inf[2] = {0, 0};
    CK(cudaMemcpyAsync(cv.data(), h->st_cost, (size_t)V * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(inf, h->st_info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (lambda >= 0) {
        CK(cudaMemcpyAsync(rel_t, h->st_rel_t, 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(poses_t, h->st_poses_t, (size_t)V * 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    double c = 0;
    for (int v = 0; v < V; ++v) c += cv[v];
    *cost = c;
    *info = inf[0] ? inf[0] : inf[1];
    return ACINO_OK;
}

int acino_lm_prepare_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t ng, const double* x_ext, const float* g,
                         const double* sw, const double* lo, const double* hi, double* gtot, uint8_t* fixed,
                         double* cost_s, void* cuda_stream) {
    DEV_ENTER("acino_lm_prepare_dev");
    if (n_frames < 0 || !x_ext || !g || !sw || !lo || !hi || !gtot || !fixed || !cost_s)
        return fail(h, ACINO_ERR_ARG, "acino_lm_prepare_dev: bad arguments");
    CK(launch_lm_prepare(n_frames, frame0, ng, x_ext, g, sw, lo, hi, gtot, fixed, cost_s, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_lm_assemble_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t ng, int n_blocks, const float* H,
                          const double* gtot, const uint8_t* fixed, const double* sw, double lambda, double* D, double* Lc,
                          double* rhs, void* cuda_stream) {
    DEV_ENTER("acino_lm_assemble_dev");
    if (n_frames < 0 || n_blocks * 3 < n_frames || !H || !gtot || !fixed || !sw || !D || !Lc || !rhs)
        return fail(h, ACINO_ERR_ARG, "acino_lm_assemble_dev: bad arguments");
    CK(launch_lm_assemble(n_frames, frame0, ng, n_blocks, H, gtot, fixed, sw, lambda, D, Lc, rhs, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_lm_step_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t ng, const double* x_ext, const double* d_ext,
                      const double* gtot, const float* H, const double* sw, const double* lo, const double* hi,
                      double* xt_ext, float* xt32, double* pred, double* step, void* cuda_stream) {
    DEV_ENTER("acino_lm_step_dev");
    if (n_frames < 0 || !x_ext || !d_ext || !gtot || !H || !sw || !lo || !hi || !xt_ext || !xt32 || !pred || !step)
        return fail(h, ACINO_ERR_ARG, "acino_lm_step_dev: bad arguments");
    CK(launch_lm_step(n_frames, frame0, ng, x_ext, d_ext, gtot, H, sw, lo, hi, xt_ext, xt32, pred, step, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_lm_reduce_dev(acino_handle* h, int n, const float* a0, const double* a1, const double* a2, const double* a3,
                        const double* m, double* out, void* cuda_stream) {
    DEV_ENTER("acino_lm_reduce_dev");
    if (n < 0 || !out) return fail(h, ACINO_ERR_ARG, "acino_lm_reduce_dev: bad arguments");
    if (!h->red_ws) {
        CK(cudaMalloc((void**)&h->red_ws, lm_reduce_ws_bytes()));
        CK(cudaMemset(h->red_ws, 0, lm_reduce_ws_bytes()));
        CK(cudaDeviceSynchronize());
    }
    // the handle is not thread-safe and calls are stream-ordered by contract: one reduction at a time uses red_ws
    CK(launch_lm_reduce(n, a0, a1, a2, a3, m, out, h->red_ws, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_bcr_factor_dev(acino_handle* h, int n_elim, const int32_t* elim, const double* D, const double* Lc, double* P,
                         double* Q, double* R, double* rhs, int32_t* info, void* cuda_stream) {
    DEV_ENTER("acino_bcr_factor_dev");
    if (n_elim < 0 || (n_elim > 0 && (!elim || !D || !Lc || !P || !Q || !R || !rhs || !info)))
        return fail(h, ACINO_ERR_ARG, "acino_bcr_factor_dev: bad arguments");
    CK(launch_bcr_factor(n_elim, elim, D, Lc, P, Q, R, rhs, info, s));
    h->launches += n_elim > 0;
    return ACINO_OK;
}

int acino_bcr_update_dev(acino_handle* h, int n_surv, const int32_t* surv, double* D, double* Lc, const double* P,
                         const double* Q, double* rhs, void* cuda_stream) {
    DEV_ENTER("acino_bcr_update_dev");
    if (n_surv < 0 || (n_surv > 0 && (!surv || !D || !Lc || !P || !Q || !rhs)))
        return fail(h, ACINO_ERR_ARG, "acino_bcr_update_dev: bad arguments");
    CK(launch_bcr_update(n_surv, surv, D, Lc, P, Q, rhs, s));
    h->launches += n_surv > 0;
    return ACINO_OK;
}

int acino_bcr_backsub_dev(acino_handle* h, int n_elim, const int32_t* elim, const double* R, const double* P,
                          const double* Q, const double* rhs, double* x, void* cuda_stream) {
    DEV_ENTER("acino_bcr_backsub_dev");
    if (n_elim < 0 || (n_elim > 0 && (!elim || !R || !P || !Q || !rhs || !x)))
        return fail(h, ACINO_ERR_ARG, "acino_bcr_backsub_dev: bad arguments");
    CK(launch_bcr_backsub(n_elim, elim, R, P, Q, rhs, x, s));
    h->launches += n_elim > 0;
    return ACINO_OK;
}

int64_t acino_sba_cam_bytes(void) { return (int64_t)sba_cam_bytes(); }

int64_t acino_sba_schur_partial_size(int n_pts, int n_cams) {
    const int64_t n = 6 * (int64_t)n_cams;
    return (int64_t)sba_schur_grid(n_pts) * (n * n + 2 * n);
}

int acino_sba_cams_dev(acino_handle* h, int n_cams, const double* params, const double* R, const double* t,
                       const double* K, const double* D, void* cams, void* cuda_stream) {
    DEV_ENTER("acino_sba_cams_dev");
    if (n_cams < 1 || n_cams > 10 || !K || !D || !cams || (!params && (!R || !t)))
        return fail(h, ACINO_ERR_ARG, "acino_sba_cams_dev: need 1..10 cameras, K, D and params or (R, t)");
    CK(launch_sba_cams(n_cams, 0, 4, params, R, t, K, D, cams, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_sba_cams_model_dev(acino_handle* h, int n_cams, int model, int n_dist, const double* params, const double* R,
                             const double* t, const double* K, const double* D, void* cams, void* cuda_stream) {
    DEV_ENTER("acino_sba_cams_model_dev");
    if (n_cams < 1 || n_cams > 10 || !K || !D || !cams || (!params && (!R || !t)))
        return fail(h, ACINO_ERR_ARG, "acino_sba_cams_model_dev: need 1..10 cameras, K, D and params or (R, t)");
    if ((model != 0 && model != 1) || n_dist < 0 || n_dist > 12 || (model == 0 && n_dist != 4))
        return fail(h, ACINO_ERR_ARG, "acino_sba_cams_model_dev: model 0 (fisheye, 4 coefficients) or 1 (standard, <= 12)");
    CK(launch_sba_cams(n_cams, model, n_dist, params, R, t, K, D, cams, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_sba_eval_dev(acino_handle* h, int n_obs, const void* cams, const double* pts, const float* uv,
                       const int32_t* cam_idx, const int32_t* pt_idx, double f_scale, double* res, double* Jc, double* Jp,
                       double* wgt, double* cost, void* cuda_stream) {
    DEV_ENTER("acino_sba_eval_dev");
    if (n_obs < 0 || !cams || !pts || !uv || !cam_idx || !pt_idx || !res || !(f_scale > 0) || (Jp && !wgt) || (Jc && !Jp))
        return fail(h, ACINO_ERR_ARG, "acino_sba_eval_dev: bad arguments");
    CK(launch_sba_eval(n_obs, cams, pts, uv, cam_idx, pt_idx, f_scale, res, Jc, Jp, wgt, cost, s));
    h->launches += n_obs > 0;
    return ACINO_OK;
}

int acino_sba_schur_dev(acino_handle* h, int n_pts, int n_cams, const int32_t* pt_ptr, const int32_t* obs,
                        const int32_t* cam_idx, const double* res, const double* Jc, const double* Jp, const double* wgt,
                        double lam, double* partial, double* S, double* rhs, void* cuda_stream) {
    DEV_ENTER("acino_sba_schur_dev");
    if (n_pts < 1 || n_cams < 1 || n_cams > 10 || !pt_ptr || !obs || !cam_idx || !res || !Jc || !Jp || !wgt || !partial || !S || !rhs)
        return fail(h, ACINO_ERR_ARG, "acino_sba_schur_dev: bad arguments");
    CK(launch_sba_schur(n_pts, n_cams, pt_ptr, obs, cam_idx, res, Jc, Jp, wgt, lam, partial, S, rhs, s));
    h->launches += 2;
    return ACINO_OK;
}

int acino_sba_dense_solve_dev(acino_handle* h, int n, double* S, double* x, int32_t* info, void* cuda_stream) {
    DEV_ENTER("acino_sba_dense_solve_dev");
    if (n < 1 || n > 96 || !S || !x || !info) return fail(h, ACINO_ERR_ARG, "acino_sba_dense_solve_dev: bad arguments");
    CK(launch_sba_dense_solve(n, S, x, info, s));
    h->launches += 1;
    return ACINO_OK;
}

int acino_sba_backsub_dev(acino_handle* h, int n_pts, int n_cams, const int32_t* pt_ptr, const int32_t* obs,
                          const int32_t* cam_idx, const double* res, const double* Jc, const double* Jp, const double* wgt,
                          double lam, const double* dc, const double* pts, double* pts_trial, double* dp,
                          void* cuda_stream) {
    DEV_ENTER("acino_sba_backsub_dev");
    if (n_pts < 0 || !pt_ptr || !obs || !cam_idx || !res || !Jp || !wgt || !pts || !pts_trial)
        return fail(h, ACINO_ERR_ARG, "acino_sba_backsub_dev: bad arguments");
    CK(launch_sba_backsub(n_pts, n_cams, pt_ptr, obs, cam_idx, res, Jc, Jp, wgt, lam, dc, pts, pts_trial, dp, s));
    h->launches += n_pts > 0;
    return ACINO_OK;
}

int acino_sba_pred_dev(acino_handle* h, int n_obs, const int32_t* cam_idx, const int32_t* pt_idx, const double* res,
                       const double* Jc, const double* Jp, const double* wgt, const double* dc, const double* dp,
                       double* pred, void* cuda_stream) {
    DEV_ENTER("acino_sba_pred_dev");
    if (n_obs < 0 || !cam_idx || !pt_idx || !res || !Jp || !wgt || !dp || !pred)
        return fail(h, ACINO_ERR_ARG, "acino_sba_pred_dev: bad arguments");
    CK(launch_sba_pred(n_obs, cam_idx, pt_idx, res, Jc, Jp, wgt, dc, dp, pred, s));
    h->launches += n_obs > 0;
    return ACINO_OK;
}

}  // extern "C"
