// TEST HARNESS (not part of the library): compiles acinoset_b200/csrc/skel_body.cuh - the exact source of the
// generic-skeleton CUDA kernels - for the host with a one-thread context, so that `pytest -m "not gpu"` can check the
// kernel arithmetic against the NumPy oracle in a container without a GPU.  Built by tests/test_skel_host.py with g++.
#include <pthread.h>

#include <cstring>
#include <map>
#include <thread>
#include <vector>

#include "../../acinoset_b200/csrc/skel_body.cuh"

using namespace acino;

struct HostCtx {
    int tid = 0, nthreads = 1;
    void sync() const {}
    void sync_part(int) const {}
};

// Multi-threaded context: T real threads and real barriers, so that the kernels' thread mappings (strides, 2-D
// decompositions, the participants-only barrier of the panel step) are exercised on the CPU as well - a CTA in slow motion.
struct MtShared {
    int nthreads;
    pthread_barrier_t all;
    std::map<int, pthread_barrier_t*> part;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    explicit MtShared(int n) : nthreads(n) { pthread_barrier_init(&all, nullptr, n); }
    pthread_barrier_t* part_barrier(int n) {
        if (n > nthreads) n = nthreads;
        pthread_mutex_lock(&mu);
        pthread_barrier_t*& b = part[n];
        if (!b) {
            b = new pthread_barrier_t;
            pthread_barrier_init(b, nullptr, n);
        }
        pthread_mutex_unlock(&mu);
        return b;
    }
};
struct MtCtx {
    int tid, nthreads;
    MtShared* sh;
    void sync() const { pthread_barrier_wait(&sh->all); }
    void sync_part(int n) const { pthread_barrier_wait(sh->part_barrier(n)); }
};
template <typename F>
static void run_mt(int nthreads, F f) {
    MtShared sh(nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] { f(MtCtx{t, nthreads, &sh}); });
    for (auto& t : th) t.join();
}

extern "C" {

int skel_host_desc_bytes() { return (int)sizeof(SkelDesc); }

// same table building as acino_skel_set (c_api.cu)
void skel_host_make_desc(SkelDesc* S, int n_parts, int n_links, int n_out, int n_cams, const int* dof_mask, const int* link_parent,
                         const int* link_flag, const double* link_tv, const unsigned long long* path, int loss_kind, double a,
                         double b, double c, double delta, const double* K, const double* D, const double* R, const double* t) {
    memset(S, 0, sizeof(*S));
    S->n_parts = n_parts; S->n_links = n_links; S->n_out = n_out; S->n_cams = n_cams;
    for (int i = 0; i < n_parts; ++i) S->dof_mask[i] = dof_mask[i] & 7;
    for (int l = 0; l < n_links; ++l) {
        S->link_parent[l] = link_parent[l];
        S->link_flag[l] = link_flag[l] & 1;
        for (int i = 0; i < 3; ++i) S->link_tv[l][i] = link_tv[3 * l + i];
    }
    for (int r = 0; r < n_out; ++r) S->path[r] = path[r];
    int k = 0;
    for (int p = 0; p < n_parts; ++p) {
        S->part_ptr[p] = k;
        for (int l = 0; l < n_links; ++l)
            if (S->link_parent[l] == p) S->part_links[k++] = l;
    }
    for (int p = n_parts; p <= SK_MAX_PARTS; ++p) S->part_ptr[p] = k;
    S->loss_kind = loss_kind; S->la = a; S->lb = b; S->lc = c; S->delta = delta;
    for (int ci = 0; ci < n_cams; ++ci) {
        SkelCam& sc = S->cam[ci];
        for (int i = 0; i < 9; ++i) sc.R[i] = R[ci * 9 + i];
        for (int i = 0; i < 3; ++i) sc.t[i] = t[ci * 3 + i];
        for (int i = 0; i < 4; ++i) sc.D[i] = D[ci * 4 + i];
        sc.fx = K[ci * 9]; sc.fy = K[ci * 9 + 4]; sc.cx = K[ci * 9 + 2]; sc.cy = K[ci * 9 + 5];
    }
}

void skel_host_eval(const SkelDesc* S, int n_frames, const double* x, const double* meas, const double* w, double* cost,
                    double* g, double* H) {
    const int P = 3 + 3 * S->n_parts;
    const size_t mo = (size_t)S->n_cams * S->n_out;
    std::vector<double> sm(SkelSmemLayout(S->n_links, S->n_out).total);
    HostCtx ctx;
    for (int n = 0; n < n_frames; ++n)
        skel_eval_frame(*S, ctx, x + (size_t)n * P, meas + n * mo * 2, w + n * mo, cost ? cost + n : nullptr,
                        g ? g + (size_t)n * P : nullptr, H ? H + (size_t)n * (P * (P + 1) / 2) : nullptr, sm.data());
}

// the same, ONE frame at a time on `nthreads` real threads with real barriers
void skel_host_eval_mt(const SkelDesc* S, int n_frames, const double* x, const double* meas, const double* w, double* cost,
                       double* g, double* H, int nthreads) {
    const int P = 3 + 3 * S->n_parts;
    const size_t mo = (size_t)S->n_cams * S->n_out;
    std::vector<double> sm(SkelSmemLayout(S->n_links, S->n_out).total);
    for (int n = 0; n < n_frames; ++n)
        run_mt(nthreads, [&](const MtCtx& ctx) {
            skel_eval_frame(*S, ctx, x + (size_t)n * P, meas + n * mo * 2, w + n * mo, cost ? cost + n : nullptr,
                            g ? g + (size_t)n * P : nullptr, H ? H + (size_t)n * (P * (P + 1) / 2) : nullptr, sm.data());
        });
}

int skel_host_band_solve_mt(long long n, int hb, int nb, double* AB, double* x, int* info, int nthreads) {
    std::vector<double> sm(band_panel_doubles(hb, nb));
    if (nb == 16) run_mt(nthreads, [&](const MtCtx& ctx) { band_cholesky_solve<16>(ctx, n, hb, AB, x, info, sm.data()); });
    else if (nb == 8) run_mt(nthreads, [&](const MtCtx& ctx) { band_cholesky_solve<8>(ctx, n, hb, AB, x, info, sm.data()); });
    else return -1;
    return 0;
}

void skel_host_prepare(int N, int P, int last_free, const double* x, const double* g, const double* sw, const double* lo,
                       const double* hi, double* gtot, unsigned char* fixed, double* cost_s) {
    skel_prepare(HostCtx(), N, P, last_free, x, g, sw, lo, hi, gtot, fixed, cost_s);
}
void skel_host_assemble(int N, int P, const double* H, const double* gtot, const unsigned char* fixed, const double* sw,
                        double lam, double* AB, double* rhs) {
    skel_assemble(HostCtx(), N, P, H, gtot, fixed, sw, lam, AB, rhs);
}
int skel_host_band_solve(long long n, int hb, int nb, double* AB, double* x, int* info) {
    std::vector<double> sm(band_panel_doubles(hb, nb));
    HostCtx ctx;
    switch (nb) {                                      // the panel width is a compile-time parameter of the kernel
        case 1: band_cholesky_solve<1>(ctx, n, hb, AB, x, info, sm.data()); return 0;
        case 4: band_cholesky_solve<4>(ctx, n, hb, AB, x, info, sm.data()); return 0;
        case 7: band_cholesky_solve<7>(ctx, n, hb, AB, x, info, sm.data()); return 0;
        case 8: band_cholesky_solve<8>(ctx, n, hb, AB, x, info, sm.data()); return 0;
        case 16: band_cholesky_solve<16>(ctx, n, hb, AB, x, info, sm.data()); return 0;
        case 32: band_cholesky_solve<32>(ctx, n, hb, AB, x, info, sm.data()); return 0;
    }
    return -1;
}
void skel_host_trial(int N, int P, int last_free, const double* x, const double* d, const double* lo, const double* hi, double* xt) {
    skel_trial(HostCtx(), N, P, last_free, x, d, lo, hi, xt);
}
void skel_host_pred(int N, int P, const double* x, const double* xt, const double* gtot, const double* H, const double* sw,
                    double* pred, double* step) {
    skel_pred(HostCtx(), N, P, x, xt, gtot, H, sw, pred, step);
}
}
