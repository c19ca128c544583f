// Internal: the library handle and the error helpers shared by the C-ABI translation units (c_api.cu, lm_plan.cu).
#pragma once
#include <string>

#include "acino_common.cuh"
#include "skel_body.cuh"
#include "stereo_body.cuh"

using namespace acino;

struct acino_handle {
    int device = 0;
    SceneF scene;
    CamD cam_d[ACINO_MAX_CAMS];
    bool have_cams = false;
    int64_t launches = 0;
    std::string err;
    // host-API staging (device)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    double* red_ws = nullptr;          // partials + ticket counter of lm_reduce (zero-initialised once)
    // tile-schedule counters of fte_eval ({ticket, finished} pairs, zero-initialised once, left zeroed by every launch): a ring,
    // so that launches of this handle that are in flight at the same time (different streams) do not share a pair
    static constexpr int kSchedSlots = 32;
    int* sched = nullptr;
    unsigned sched_next = 0;
    int* next_sched() { return sched ? sched + 2 * (sched_next++ % kSchedSlots) : nullptr; }
    cudaStream_t stream = nullptr;
    // host-API pipeline: H2D / compute / D2H on three streams, chunked, chained with events
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    static constexpr int kMaxChunks = 64;
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {};
    bool pipe_ready = false;
    // generic-skeleton variant (build.py): host copy + device copy of the descriptor
    SkelDesc skel;
    SkelDesc* d_skel = nullptr;
    bool have_skel = false;
    // pairwise extrinsic calibration (stereo) problem: device buffers owned by the handle
    StereoCam st_c1, st_c2;
    int st_V = 0, st_M = 0;
    double* st_buf = nullptr;          // obj | img1 | img2 | rel | rel_t | poses | poses_t | cost | S | back | d_rel | info
    double *st_obj = nullptr, *st_img1 = nullptr, *st_img2 = nullptr, *st_rel = nullptr, *st_rel_t = nullptr, *st_poses = nullptr,
           *st_poses_t = nullptr, *st_cost = nullptr, *st_S = nullptr, *st_back = nullptr, *st_drel = nullptr;
    int* st_info = nullptr;
};

inline thread_local std::string g_err;

inline int fail(acino_handle* h, int code, const std::string& msg) {
    g_err = msg;
    if (h) h->err = msg;
    return code;
}
inline int cuda_fail(acino_handle* h, cudaError_t e, const char* what) {
    return fail(h, ACINO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return cuda_fail(h, _e, #call);       \
    } while (0)

