// TEST HARNESS (not part of the library): acinoset_b200/csrc/stereo_body.cuh - the source of the pairwise extrinsic
// calibration kernels - compiled for the host with a one-thread context, mirroring acino_stereo_set / _init / _step of
// csrc/c_api.cu so that acinoset_b200/stereo.py's accept / reject loop can be exercised without a GPU.
#include <vector>

#include "../../acinoset_b200/csrc/stereo_body.cuh"

using namespace acino;

struct HostCtx {
    int tid = 0, nthreads = 1;
};

static StereoCam g_c1, g_c2;
static int g_V = 0, g_M = 0;
static std::vector<double> g_obj, g_img1, g_img2;

static StereoCam make_cam(const double* K, const double* D, int nd, int model) {
    StereoCam c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    for (int i = 0; i < 12; ++i) c.D[i] = i < nd ? D[i] : 0.0;
    c.model = model;
    return c;
}

extern "C" {

int stereo_host_set(int V, int M, const double* obj, const double* img1, const double* img2, const double* K1, const double* D1,
                    int nd1, const double* K2, const double* D2, int nd2, int model) {
    g_V = V; g_M = M;
    g_obj.assign(obj, obj + (size_t)M * 3);
    g_img1.assign(img1, img1 + (size_t)V * M * 2);
    g_img2.assign(img2, img2 + (size_t)V * M * 2);
    g_c1 = make_cam(K1, D1, nd1, model);
    g_c2 = make_cam(K2, D2, nd2, model);
    return 0;
}

int stereo_host_init(double* poses, double* cost) {
    stereo_init_poses(HostCtx(), g_c1, g_c2, g_V, g_M, g_obj.data(), g_img1.data(), g_img2.data(), poses, cost);
    return 0;
}

int stereo_host_step(const double* rel, const double* poses, double lambda, double* rel_t, double* poses_t, double* cost, int* info) {
    const int V = g_V, M = g_M;
    std::vector<double> cv(V), S((size_t)V * 42), back((size_t)V * 42);
    double d_rel[6];
    int inf[2] = {0, 0};
    const double* rel_eval = rel;
    const double* poses_eval = poses;
    if (lambda >= 0) {
        stereo_view_blocks(HostCtx(), g_c1, g_c2, V, M, g_obj.data(), g_img1.data(), g_img2.data(), rel, poses, lambda, 1, cv.data(),
                           S.data(), back.data(), inf);
        stereo_reduce_solve(HostCtx(), V, S.data(), d_rel, inf + 1);
        stereo_update(HostCtx(), V, rel, poses, back.data(), d_rel, rel_t, poses_t);
        rel_eval = rel_t;
        poses_eval = poses_t;
    }
    stereo_view_blocks(HostCtx(), g_c1, g_c2, V, M, g_obj.data(), g_img1.data(), g_img2.data(), rel_eval, poses_eval, 0.0, 0, cv.data(),
                       nullptr, nullptr, nullptr);
    double c = 0;
    for (int v = 0; v < V; ++v) c += cv[v];
    *cost = c;
    *info = inf[0] ? inf[0] : inf[1];
    return 0;
}
}
