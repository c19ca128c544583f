// Pairwise fisheye extrinsic calibration kernels (reference calib.py:125-134 = cv2.fisheye.stereoCalibrate with fixed
// intrinsics).  Arithmetic in stereo_body.cuh (shared with the CPU test harness).
//   stereo_init    thread per (view, camera): homography -> board pose -> damped Gauss-Newton refinement
//   stereo_blocks  thread per view: residuals + Jacobians of both cameras, 6 x 6 view block eliminated (Schur term)
//   stereo_solve   one thread: fixed-order sum of the Schur terms, 6 x 6 Cholesky -> relative-pose step
//   stereo_update  thread per view (+ one for the relative pose): multiplicative rotation update -> trial point
#include <cuda_runtime.h>

#include "stereo_body.cuh"

namespace acino {

struct StGridCtx {
    int tid, nthreads;
};

__global__ void stereo_init_kernel(StereoCam c1, StereoCam c2, int V, int M, const double* obj, const double* img1,
                                   const double* img2, double* pose_out, double* cost_out) {
    const StGridCtx ctx{(int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x)};
    stereo_init_poses(ctx, c1, c2, V, M, obj, img1, img2, pose_out, cost_out);
}
__global__ void stereo_blocks_kernel(StereoCam c1, StereoCam c2, int V, int M, const double* obj, const double* img1,
                                     const double* img2, const double* rel, const double* poses, double lam, int blocks,
                                     double* out_cost, double* out_S, double* out_back, int* out_info) {
    const StGridCtx ctx{(int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x)};
    stereo_view_blocks(ctx, c1, c2, V, M, obj, img1, img2, rel, poses, lam, blocks, out_cost, out_S, out_back, out_info);
}
__global__ void stereo_solve_kernel(int V, const double* S_all, double* d_rel, int* info) {
    const StGridCtx ctx{(int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x)};
    stereo_reduce_solve(ctx, V, S_all, d_rel, info);
}
__global__ void stereo_update_kernel(int V, const double* rel, const double* poses, const double* back, const double* d_rel,
                                     double* rel_t, double* poses_t) {
    const StGridCtx ctx{(int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x)};
    stereo_update(ctx, V, rel, poses, back, d_rel, rel_t, poses_t);
}

static inline int st_grid(int n, int b) { return n < 1 ? 1 : (n + b - 1) / b; }

cudaError_t launch_stereo_init(const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj, const double* img1,
                               const double* img2, double* pose_out, double* cost_out, cudaStream_t s) {
    stereo_init_kernel<<<st_grid(2 * V, 32), 32, 0, s>>>(c1, c2, V, M, obj, img1, img2, pose_out, cost_out);
    return cudaGetLastError();
}
cudaError_t launch_stereo_blocks(const StereoCam& c1, const StereoCam& c2, int V, int M, const double* obj, const double* img1,
                                 const double* img2, const double* rel, const double* poses, double lam, int blocks,
                                 double* out_cost, double* out_S, double* out_back, int* out_info, cudaStream_t s) {
    stereo_blocks_kernel<<<st_grid(V, 32), 32, 0, s>>>(c1, c2, V, M, obj, img1, img2, rel, poses, lam, blocks, out_cost, out_S,
                                                       out_back, out_info);
    return cudaGetLastError();
}
cudaError_t launch_stereo_solve(int V, const double* S_all, double* d_rel, int* info, cudaStream_t s) {
    stereo_solve_kernel<<<1, 32, 0, s>>>(V, S_all, d_rel, info);
    return cudaGetLastError();
}
cudaError_t launch_stereo_update(int V, const double* rel, const double* poses, const double* back, const double* d_rel,
                                 double* rel_t, double* poses_t, cudaStream_t s) {
    stereo_update_kernel<<<st_grid(V + 1, 32), 32, 0, s>>>(V, rel, poses, back, d_rel, rel_t, poses_t);
    return cudaGetLastError();
}

}  // namespace acino
