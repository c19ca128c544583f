/*
 * acino_b200.h - C ABI of libacino_b200.so: the B200 (sm_100a) implementation of
 * AcinoSet's reprojection / trajectory-optimisation hot path.
 *
 * The reference (African-Robotics-Unit/AcinoSet) has no FFI of its own - its boundary is a
 * set of Python callables.  Each entry point below names the reference
 * function(s) whose arithmetic it replaces; the ctypes binding that puts these behind the
 * reference's own names lives in acinoset_b200/_lib.py and is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error (acino_last_error() has the text);
 *     nothing throws across the ABI;
 *   - "_dev" entry points take DEVICE pointers and are stream-ordered (no hidden syncs);
 *     entry points without the suffix take HOST pointers, copy in/out and synchronise;
 *   - all arrays are dense row-major; the caller owns every buffer; the library owns only
 *     the handle (camera table, loss constants, workspace);
 *   - one handle per GPU; a handle is not thread-safe, different handles are independent;
 *   - state vectors use the 25 "active" pose slots in the order of the reference's saved
 *     result pickle (all_optimizations.py:540-556):
 *       x,y,z, phi0,phi1,phi3, theta0..theta13, psi0,psi1,psi3,psi4,psi5.
 */
#ifndef ACINO_B200_H
#define ACINO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACINO_N_ACTIVE   25    /* active pose parameters per frame                      */
#define ACINO_N_MARKERS  20    /* cheetah markers, order of all_optimizations.py:170-178 */
#define ACINO_N_UPPER    325   /* packed upper triangle of a 25x25 block, row-major      */
#define ACINO_MAX_CAMS   16

#define ACINO_OK              0
#define ACINO_ERR_ARG        -1
#define ACINO_ERR_CUDA       -2
#define ACINO_ERR_STATE      -3

typedef struct acino_handle acino_handle;

/* ---- lifetime ------------------------------------------------------------------------- */
int acino_create(acino_handle** out, int device);
int acino_destroy(acino_handle* h);
const char* acino_last_error(const acino_handle* h);   /* h may be NULL: last global error */
int acino_version(void);
/* number of kernel launches issued through this handle so far (bench.py "gpu_launches") */
int64_t acino_launch_count(const acino_handle* h);

/* ---- scene ---------------------------------------------------------------------------- */
/* Camera table: K [C][3][3], D [C][4], R [C][3][3] (world->camera), t [C][3]; the arrays
 * utils.load_scene returns (src/calib/utils.py:84-101), D reshaped (-1,4) as at
 * all_optimizations.py:221. */
int acino_set_cameras(acino_handle* h, int n_cams, const double* K, const double* D,
                      const double* R, const double* t);
/* Redescending-loss break points (a,b,c); default (3,10,20), all_optimizations.py:25-27. */
int acino_set_redescending(acino_handle* h, double a, double b, double c);

/* ---- FTE: residual + Jacobian evaluation (the north-star kernel) ------------------------
 * Replaces, per frame n: pose_constraint (all_optimizations.py:359-365), measurement_constraints
 * (:394-399) through pt3d_to_2d (:193-209), the measurement term of obj (:494-497) with
 * misc.redescending_loss (build.py:382-395), and the automatic differentiation IPOPT's ASL
 * does on them.
 *   x    [N][25]        active pose state
 *   meas [N][C][20][2]  measured pixels (u,v)
 *   w    [N][C][20]     meas_err_weight (1/R if likelihood > thresh else 0, :302-308)
 *   cost [N]            sum_{c,l,d} rho(w * r)
 *   g    [N][25]        d cost / d x_n
 *   H    [N][325]       packed upper triangle of sum psi(w r) w^2 J^T J  (psi = max(rho'/e, 1-sigma_a))
 * Any of cost/g/H may be NULL (not written).  meas must be 8-byte aligned (ACINO_ERR_ARG otherwise); with every
 * pointer 16-byte aligned the tiles move by bulk async copies (TMA), otherwise by plain loads / stores - same bits.
 * Launches of more than one wave of 8-frame tiles draw their tiles from a device counter pair taken from a ring of 32 that
 * the handle owns (results do not depend on the schedule): at most 32 such launches of ONE handle may be in flight at the
 * same time (different streams, or captured into graphs that run concurrently); use one handle per stream beyond that. */
int acino_fte_eval_dev(acino_handle* h, int n_frames, const float* x, const float* meas,
                       const float* w, float* cost, float* g, float* H, void* cuda_stream);
int acino_fte_eval(acino_handle* h, int n_frames, const float* x, const float* meas,
                   const float* w, float* cost, float* g, float* H);
/* Name (as ncu / the launch list shows it) of the kernel acino_fte_eval[_dev] launches for a batch of
 * n_frames frames: bench.py's roofline.kernel. */
const char* acino_fte_eval_kernel_name(int n_frames);

/* Reprojection only: pose_to_3d (all_optimizations.py:186) + project_points_fisheye for every
 * camera (calib.py:132-136; save_3d_cheetah_as_2d at all_optimizations.py:560).
 *   pos [N][20][3] world marker positions (may be NULL), uv [N][C][20][2] pixels (may be NULL) */
int acino_fk_project_dev(acino_handle* h, int n_frames, const float* x, float* pos, float* uv,
                         void* cuda_stream);
int acino_fk_project(acino_handle* h, int n_frames, const float* x, float* pos, float* uv);

/* Dense measurement Jacobian (the EKF's linearisation): h_function (all_optimizations.py:615-621) and
 * the analytic replacement of numerical_jacobian (:634-649, used at :800-806) for every camera at once.
 *   uv [N][C][20][2]      predicted pixels h(x)                       (may be NULL)
 *   J  [N][C][20][2][25]  d uv / d x in the active-slot order above   (may be NULL)
 * acinoset_b200/ekf.py permutes the columns into the EKF's joint-grouped state order (:734-746). */
int acino_fte_jac_dev(acino_handle* h, int n_frames, const float* x, float* uv, float* J,
                      void* cuda_stream);
int acino_fte_jac(acino_handle* h, int n_frames, const float* x, float* uv, float* J);

/* ---- camera geometry (fp64, host pointers) -------------------------------------------------
 * Single-camera arguments: K [3][3], D [4], R [3][3] (world->camera), t [3]. */

/* project_points_fisheye(obj_pts, k, d, r, t) (calib.py:132-136): X [n][3] -> uv [n][2].
 * No skew, no behind-camera guard, exactly like cv2.fisheye.projectPoints / pt3d_to_2d. */
int acino_project_points(acino_handle* h, int n, const double* X, const double* K, const double* D,
                         const double* R, const double* t, double* uv);

/* cv2.fisheye.undistortPoints(pts, K, D) as called at calib.py:124-125 (no R/P, default criteria):
 * uv [n][2] pixels -> xn [n][2] normalised coordinates; non-converged points come back as -1e6. */
int acino_undistort_points(acino_handle* h, int n, const double* uv, const double* K, const double* D,
                           double* xn);

/* triangulate_points_fisheye(img_pts_1, img_pts_2, k1,d1,r1,t1, k2,d2,r2,t2) (calib.py:121-130):
 * uv1, uv2 [n][2] -> X [n][3]. */
int acino_triangulate_points(acino_handle* h, int n, const double* uv1, const double* uv2,
                             const double* K1, const double* D1, const double* R1, const double* t1,
                             const double* K2, const double* D2, const double* R2, const double* t2,
                             double* X);

/* Pinhole twins of the three calls above: project_points (calib.py:64-66, cv2.projectPoints),
 * create_undistort_point_function (:25-30, cv2.undistortPoints with P = K: to_pixels = 1) and
 * triangulate_points (:52-61, cv2.undistortPoints x2 + cv2.triangulatePoints).  dist holds n_dist <= 14
 * coefficients in OpenCV order [k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tauX tauY] (the reference calibrates the
 * 8-coefficient rational model, calib.py:18); missing ones are 0, non-zero tilt terms are ACINO_ERR_ARG.
 * Undistortion = OpenCV's default criteria: exactly 5 fixed-point iterations. */
int acino_project_points_pinhole(acino_handle* h, int n, const double* X, const double* K,
                                 const double* dist, int n_dist, const double* R, const double* t,
                                 double* uv);
int acino_undistort_points_pinhole(acino_handle* h, int n, const double* uv, const double* K,
                                   const double* dist, int n_dist, int to_pixels, double* out);
int acino_triangulate_points_pinhole(acino_handle* h, int n, const double* uv1, const double* uv2,
                                     const double* K1, const double* dist1, int n_dist1,
                                     const double* R1, const double* t1, const double* K2,
                                     const double* dist2, int n_dist2, const double* R2,
                                     const double* t2, double* X);

/* get_pairwise_3d_points_from_df (calib.py:394-423) on dense tensors, cameras from
 * acino_set_cameras: uv [N][C][L][2], valid [N][C][L] (1 = the row survives the caller's likelihood
 * filter) -> pos [N][L][3] = unweighted mean of the adjacent-pair (c, c+1) triangulations in the
 * order 0-1, 1-2, ... (NaN where no pair saw the point), count [N][L] (may be NULL). */
int acino_triangulate_pairwise(acino_handle* h, int n_frames, int n_markers, const double* uv,
                               const uint8_t* valid, double* pos, int32_t* count);

/* Generic skeleton-pickle forward kinematics: the pose_to_3d that build_model lambdifies
 * (src/build.py:32-95), quirks included.  The host (acinoset_b200/skeleton.py) flattens the skeleton:
 * dof_mask [n_parts] (bit0 phi/x, bit1 theta/y, bit2 psi/z), and per link in order parent / child part
 * index, flag (1 = the parent's local rotation is used transposed, 0 = as is) and rest-pose offset
 * tv [n_links][3].  x [N][3 + 3 n_parts] = [x,y,z,*phi,*theta,*psi] -> pos [N][n_parts][3] (fp64,
 * host pointers).  n_parts <= 32. */
int acino_generic_fk(acino_handle* h, int n_frames, int n_parts, int n_links, const int32_t* dof_mask,
                     const int32_t* link_parent, const int32_t* link_child, const int32_t* link_flags,
                     const double* link_tv, const double* x, double* pos);

/* ---- generic-skeleton FTE: the data-driven variant of the reference (src/build.py:28-332), fp64 ---------------
 * Any skeleton pickle {links, dofs, positions, markers}; state per frame x [P], P = 3 + 3 n_parts =
 * [x,y,z, phi_0.., theta_0.., psi_0..] (build.py:47-49,88); output rows = parts in pose_dict order (:82-86), which
 * the reference pairs index-by-index with the measured markers (:189-198,276-285).
 * acino_skel_set: the flattened builder tables of acinoset_b200/skeleton.py (see acino_generic_fk) plus, per output
 * row, path = bit mask of the links whose increments sum to its pose; cameras come from acino_set_cameras (call it
 * first).  loss_kind 0: redescending_loss(|w r|, a, b, c) (all_optimizations.py:497); 1: |w r| (build.py:299) with
 * Gauss-Newton curvature weight 1 / max(|w r|, delta).  n_parts, n_out <= 32, n_links <= 40. */
int acino_skel_set(acino_handle* h, int n_parts, int n_links, int n_out, const int32_t* dof_mask,
                   const int32_t* link_parent, const int32_t* link_flag, const double* link_tv,
                   const uint64_t* path, int loss_kind, double a, double b, double c, double delta);
/* Replaces pose_constraint + measurement_constraints + the measurement term of obj (build.py:219-224,276-285,
 * 294-300) and their differentiation:  x [N][P], meas [N][C][n_out][2], w [N][C][n_out] (meas_err_weight,
 * :166-172; 0 = skipped) -> cost [N], g [N][P], H [N][P(P+1)/2] packed upper triangle of sum psi w^2 J^T J.
 * Any output may be NULL.  _dev: device pointers, stream-ordered; the other: host pointers. */
int acino_skel_eval_dev(acino_handle* h, int n_frames, const double* x, const double* meas,
                        const double* w, double* cost, double* g, double* H, void* cuda_stream);
int acino_skel_eval(acino_handle* h, int n_frames, const double* x, const double* meas, const double* w,
                    double* cost, double* g, double* H);
/* Levenberg-Marquardt building blocks replacing opt.solve (build.py:306-332) on
 *   F = sum rho(w r) + sum_{n>=3,p} q_p (third difference / h^2)^2     (backwards_euler_* + constant_acc, :231-261)
 * sw [P] = 2 q_p / h^4 (q_p = 0.002, :173-177), lo / hi [P] box bounds (:263-266), last_free = 1: the last frame is
 * unbounded like the reference's range(1, N).  fixed [N][P]: variables frozen at a bound; cost_s [N] smoothness cost.
 * AB [N P][3P + 1] row-wise lower band of (B + lambda diag B), AB[i][k] = B[i][i-k]; rhs [N P] = -gradient. */
int acino_skel_prepare_dev(acino_handle* h, int n_frames, int last_free, const double* x, const double* g,
                           const double* sw, const double* lo, const double* hi, double* gtot,
                           uint8_t* fixed, double* cost_s, void* cuda_stream);
int acino_skel_assemble_dev(acino_handle* h, int n_frames, const double* H, const double* gtot,
                            const uint8_t* fixed, const double* sw, double lambda, double* AB,
                            double* rhs, void* cuda_stream);
/* in-place band Cholesky of AB [n][half_bandwidth + 1] and solve: x (rhs in, solution out); info != 0: index + 1
 * of the first non-positive pivot (AB, x then undefined) */
int acino_band_solve_dev(acino_handle* h, int64_t n, int half_bandwidth, double* AB, double* x,
                         int32_t* info, void* cuda_stream);
int acino_skel_trial_dev(acino_handle* h, int n_frames, int last_free, const double* x, const double* d,
                         const double* lo, const double* hi, double* xt, void* cuda_stream);
/* pred [N]: decrease of the quadratic model for s = xt - x; step [N] = max |s| per frame */
int acino_skel_pred_dev(acino_handle* h, int n_frames, const double* x, const double* xt,
                        const double* gtot, const double* H, const double* sw, double* pred,
                        double* step, void* cuda_stream);

/* ---- pairwise extrinsic calibration of two fisheye cameras, fp64, host pointers -------------------
 * Replaces calibrate_pair_extrinsics_fisheye (calib.py:125-134: cv2.fisheye.stereoCalibrate, CALIB_FIX_INTRINSIC):
 * relative pose (R, T) of camera 2 w.r.t. camera 1 from n_views checkerboard views seen by both, by minimising the
 * reprojection error in both cameras over (R, T) and one board pose per view.  obj [M][3] board corners (z = 0),
 * img1 / img2 [V][M][2] pixels, K [3][3], D [4].  Poses are 12 doubles: R (row-major 3 x 3) then t.
 * acino_stereo_set    uploads the problem (device buffers owned by the handle);
 * acino_stereo_init   per view and camera: board pose from the planar homography + damped Gauss-Newton refinement:
 *                     poses [V][2][12], cost [V][2] (sum of squared pixel errors; < 0 = failed);
 * acino_stereo_step   lambda < 0: cost of (rel [12], poses [V][12] = board poses in camera 1) only; lambda >= 0: one
 *                     Levenberg-Marquardt step (per-view blocks eliminated, rotations updated as R exp([d]x)) ->
 *                     trial point rel_t, poses_t and ITS cost.  info != 0: a block was not positive definite.
 * The accept / reject loop is driven from acinoset_b200/stereo.py. */
int acino_stereo_set(acino_handle* h, int n_views, int n_points, const double* obj, const double* img1,
                     const double* img2, const double* K1, const double* D1, const double* K2,
                     const double* D2);
/* the same problem for the standard camera model (calibrate_pair_extrinsics, calib.py:41-49: cv2.stereoCalibrate with
 * CALIB_FIX_INTRINSIC); dist as in acino_project_points_pinhole */
int acino_stereo_set_pinhole(acino_handle* h, int n_views, int n_points, const double* obj,
                             const double* img1, const double* img2, const double* K1,
                             const double* dist1, int n_dist1, const double* K2, const double* dist2,
                             int n_dist2);
int acino_stereo_init(acino_handle* h, double* poses, double* cost);
int acino_stereo_step(acino_handle* h, const double* rel, const double* poses, double lambda,
                      double* rel_t, double* poses_t, double* cost, int32_t* info);

/* ---- FTE solve building blocks (device pointers, stream-ordered) ------------------------------
 * Together they replace `opt.solve(m)` (all_optimizations.py:503-524): a projected
 * Levenberg-Marquardt loop on  F(x) = sum rho(w r) + sum_{n>=3,p} q_p (third difference / Ts^2)^2
 * (backwards_euler_pos/_vel + constant_acc, :369-391; weights 1/Q, :245-252,310-315) with the 21
 * pose bounds of :403-483.  The loop itself (and the one all_gather per iteration when frames are
 * sharded over GPUs) is driven from acinoset_b200/fte.py.
 *
 * Frames are addressed globally: this rank holds global frames [frame0, frame0 + n_frames) of
 * n_frames_global; x_ext / d_ext are [(n_frames + 6)][25] fp64 with 3 halo frames on each side.
 * sw[25] = 2 q_p / Ts^4, lo/hi[25] the box bounds (+-inf where free).  Super-blocks hold 3 frames
 * (75 unknowns): D, Lc, P, Q are [M][75][75] fp64 row-major, rhs / x [M][75]. */
int acino_lm_prepare_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t n_frames_global,
                         const double* x_ext, const float* g, const double* sw, const double* lo,
                         const double* hi, double* gtot, uint8_t* fixed, double* cost_s,
                         void* cuda_stream);
int acino_lm_assemble_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t n_frames_global,
                          int n_blocks, const float* H, const double* gtot, const uint8_t* fixed,
                          const double* sw, double lambda, double* D, double* Lc, double* rhs,
                          void* cuda_stream);
int acino_lm_step_dev(acino_handle* h, int n_frames, int64_t frame0, int64_t n_frames_global,
                      const double* x_ext, const double* d_ext, const double* gtot, const float* H,
                      const double* sw, const double* lo, const double* hi, double* xt_ext,
                      float* xt32, double* pred, double* step, void* cuda_stream);
/* out[0..3] = fixed-order fp64 sums of a0 (float), a1, a2, a3; out[4] = max of m (NULL = skipped) */
int acino_lm_reduce_dev(acino_handle* h, int n, const float* a0, const double* a1, const double* a2,
                        const double* a3, const double* m, double* out, void* cuda_stream);

/* Block cyclic reduction of a block-tridiagonal SPD chain (schedule: acinoset_b200/bcr.py).
 * elim / surv are [n][3] int32 rows (block, left, right) / (block, eliminated-left, eliminated-right),
 * -1 = none.  factor writes the factor of D_e, R = L Delta^1/2 (strict lower triangle L, diagonal Delta^-1/2), to
 * R[e] (D_e itself is left alone: small levels run two CTAs per block and the other one may still be reading it)
 * and overwrites rhs_e with z = R^-1 b; info != 0 flags a non-positive pivot (block index + 1). */
int acino_bcr_factor_dev(acino_handle* h, int n_elim, const int32_t* elim, const double* D, const double* Lc,
                         double* P, double* Q, double* R, double* rhs, int32_t* info, void* cuda_stream);
int acino_bcr_update_dev(acino_handle* h, int n_surv, const int32_t* surv, double* D, double* Lc,
                         const double* P, const double* Q, double* rhs, void* cuda_stream);
int acino_bcr_backsub_dev(acino_handle* h, int n_elim, const int32_t* elim, const double* R,
                          const double* P, const double* Q, const double* rhs, double* x,
                          void* cuda_stream);

/* ---- FTE solve: the whole Levenberg-Marquardt attempt as stream-ordered phases --------------------
 * Replaces `results = opt.solve(m, tee=True)` (all_optimizations.py:503-524) without a host in the loop:
 * lambda, the objective, the accept / reject decision and the convergence test live in a device-resident
 * control block; a caller enqueues the phases of one attempt (or replays a CUDA graph of them) and reads
 * the pinned mirror of the control block when it wants to know whether the solve has finished.
 * Multi-GPU (frames sharded over ranks): the caller performs the two exchanges between the phases -
 * all_gather(payload -> gathered) after REDUCE and all_gather(sums_local -> sums_all) after TRIAL; with
 * world == 1 pass gathered = payload and sums_all = sums_local (no exchange).
 * All pointers are DEVICE pointers unless stated; the caller owns every buffer for the plan's lifetime. */
#define ACINO_LM_PAYLOAD   22800   /* doubles: D_first, D_last, Lc_first, Lc_last (75x75 each), rhs_first, rhs_last,
                                      frozen_first, frozen_last (75 each) */
#define ACINO_LM_SUMS      8       /* doubles per rank: sum cost, sum smoothness cost, sum model reduction, -, max |step|, - */
#define ACINO_LM_CTL       32      /* doubles: see acinoset_b200/lm.py CTL_* (lambda, F, ..., done, status) */
#define ACINO_LM_HIST      8       /* doubles per logged attempt: F, F_trial, lambda, rho, |step|inf, accepted, pred, - */
typedef struct acino_lm_desc {
    int32_t n_frames;              /* N: frames of this rank's shard */
    int32_t n_blocks;              /* M = ceil(N / 3) super-blocks of 3 frames x 25 parameters */
    int32_t rank, world;
    int64_t frame0, n_global;      /* first global frame of the shard, global frame count */
    const float* meas;             /* [N][C][20][2] */
    const float* w;                /* [N][C][20] */
    const double* sw;              /* [25] 2 q_p / Ts^4 (all_optimizations.py:245-252, 369-391) */
    const double* lo;              /* [25] pose bounds (:403-483), +-inf where free */
    const double* hi;
    double* x_ext[2];              /* [N + 6][25] accepted / trial state, 3 halo frames per side */
    float* x32[2];                 /* [N][25] fp32 copy the evaluation consumes */
    float* cost[2];                /* [N]      outputs of fte_eval */
    float* g[2];                   /* [N][25] */
    float* H[2];                   /* [N][325] */
    double* gtot[2];               /* [N][25] total gradient (measurement + smoothness) */
    uint8_t* fixed[2];             /* [N][25] frozen (bound-active) variables */
    double* cost_s[2];             /* [N] smoothness cost per frame */
    double* pred;                  /* [N] model reduction per frame */
    double* step;                  /* [N] max |step| per frame */
    double* D;                     /* [M][75][75] */
    double* Lc;                    /* [M][75][75] */
    double* P;                     /* [M][75][75] */
    double* Q;                     /* [M][75][75] */
    double* R;                     /* [M][75][75] factors of the blocks the dense levels eliminate */
    double* rhs;                   /* [M][75] */
    double* dx;                    /* [M][75] the step */
    double* dhalo;                 /* [2][75] step of the neighbouring ranks' interface blocks (0 at the global ends) */
    int32_t* info;                 /* [1] != 0: non-positive pivot (block index + 1) */
    int32_t n_elim0, n_surv0;      /* structured level 0 (csrc/lm_l0.cu): eliminated / surviving blocks */
    const int32_t* elim0;          /* [n_elim0][3] (block, left, right) */
    const int32_t* surv0;          /* [n_surv0][3] (block, eliminated left or -1, eliminated right or -1) */
    int32_t n_levels;              /* dense levels below (csrc/bcr.cu) */
    const int32_t* level_counts;   /* HOST [n_levels][2] (n_elim, n_surv) */
    const int32_t* sched;          /* per level: n_elim elim rows then n_surv surv rows, [.][3] each */
    double* payload;               /* [ACINO_LM_PAYLOAD] this rank's reduced end blocks (world > 1) */
    const double* gathered;        /* [world][ACINO_LM_PAYLOAD] */
    double* cD; double* cLc; double* cP; double* cQ; double* cR;   /* interface chain [2 world][75][75] */
    double* crhs; double* cx;      /* [2 world][75] */
    int32_t n_clevels;
    const int32_t* clevel_counts;  /* HOST */
    const int32_t* csched;
    double* sums_local;            /* [ACINO_LM_SUMS] */
    const double* sums_all;        /* [world][ACINO_LM_SUMS] */
    double* ctl;                   /* [ACINO_LM_CTL] */
    double* ctl_host;              /* HOST, pinned: mirror written at the end of INIT_FINISH and DECIDE */
    double* hist;                  /* [hist_cap][ACINO_LM_HIST] */
    int32_t hist_cap;
} acino_lm_desc;
typedef struct acino_lm_plan acino_lm_plan;
enum {
    ACINO_LM_INIT_EVAL = 0,        /* evaluate the accepted state: fte_eval, gradient / frozen set, local sums */
    ACINO_LM_INIT_FINISH = 1,      /* F <- sum over ranks */
    ACINO_LM_REDUCE = 2,           /* assemble (B + lam diag B), eliminate down to the end blocks (world > 1: payload) */
    ACINO_LM_BACKSUB = 3,          /* world > 1: interface chain from `gathered`; back-substitution -> dx, dhalo */
    ACINO_LM_TRIAL = 4,            /* projected trial point, fte_eval there, gradient / frozen set, local sums */
    ACINO_LM_DECIDE = 5            /* gain ratio, lambda update, commit trial -> accepted, convergence test */
};
int acino_lm_desc_size(void);      /* sizeof(acino_lm_desc): lets a foreign-language binding check its struct layout */
int acino_lm_plan_create(acino_handle* h, const acino_lm_desc* desc, acino_lm_plan** out);
int acino_lm_plan_destroy(acino_lm_plan* plan);
int acino_lm_enqueue(acino_handle* h, acino_lm_plan* plan, int phase, void* cuda_stream);

/* ---- sparse bundle adjustment building blocks (fp64, device pointers, stream-ordered) ---------
 * Replace cost_func_points_extrinsics / cost_func_points_only (calib.py:312-316,355-359) and the
 * finite-difference Jacobian + TRF step SciPy's least_squares performs on them (calib.py:335,381).
 * Parameter layout = the reference's (calib.py:345-352, 373-375):
 *   params = [rvec_0..rvec_{C-1} | t_0..t_{C-1}] (6C doubles), points [n_pts][3] separately.
 * Observation i: pixel uv[i] (float32 like points_2d, calib.py:259), camera cam_idx[i], point
 * pt_idx[i]; residual order [u0,v0,u1,v1,...].  Per observation: res [2], Jc [2][6] (d/d rvec,
 * d/d t of its camera), Jp [2][3] (d/d its point), wgt [2] = Cauchy rho'((f/f_scale)^2),
 * cost = 0.5 f_scale^2 ln(1 + (f/f_scale)^2) summed over the two coordinates (SciPy's loss). */
int64_t acino_sba_cam_bytes(void);                 /* size of one opaque device camera record */
/* params != NULL: cameras from the parameter vector (Rodrigues + derivative); else from R, t */
int acino_sba_cams_dev(acino_handle* h, int n_cams, const double* params, const double* R,
                       const double* t, const double* K, const double* D, void* cams,
                       void* cuda_stream);
/* Same with an explicit camera model: 0 = fisheye (n_dist = 4), 1 = OpenCV's standard model with n_dist <= 12 coefficients
 * [k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4] per camera (cv2.projectPoints, calib.py:64-66: what app.sba_board_points passes,
 * app.py:215-218).  D is [n_cams][n_dist]. */
int acino_sba_cams_model_dev(acino_handle* h, int n_cams, int model, int n_dist, const double* params,
                             const double* R, const double* t, const double* K, const double* D, void* cams,
                             void* cuda_stream);
/* Jp == NULL: residuals (+ cost) only; Jc == NULL: points-only problem */
int acino_sba_eval_dev(acino_handle* h, int n_obs, const void* cams, const double* pts,
                       const float* uv, const int32_t* cam_idx, const int32_t* pt_idx, double f_scale,
                       double* res, double* Jc, double* Jp, double* wgt, double* cost,
                       void* cuda_stream);
/* Schur complement onto the 6C x 6C camera system with Marquardt damping lam; points in CSR form
 * (pt_ptr [n_pts+1], obs [n_obs] observation ids grouped by point).  partial: workspace of
 * acino_sba_schur_partial_size(n_pts, n_cams) doubles.  S [6C][6C], rhs [6C]. */
int64_t acino_sba_schur_partial_size(int n_pts, int n_cams);
int acino_sba_schur_dev(acino_handle* h, int n_pts, int n_cams, const int32_t* pt_ptr,
                        const int32_t* obs, const int32_t* cam_idx, const double* res, const double* Jc,
                        const double* Jp, const double* wgt, double lam, double* partial, double* S,
                        double* rhs, void* cuda_stream);
/* x <- S^-1 x, n <= 96 (Cholesky in one CTA); info != 0: non-positive pivot */
int acino_sba_dense_solve_dev(acino_handle* h, int n, double* S, double* x, int32_t* info,
                              void* cuda_stream);
/* per point: dp = -(V + lam diag V)^-1 (gv + sum W^T dc); pts_trial = pts + dp (dc / Jc may be NULL) */
int acino_sba_backsub_dev(acino_handle* h, int n_pts, int n_cams, const int32_t* pt_ptr,
                          const int32_t* obs, const int32_t* cam_idx, const double* res, const double* Jc,
                          const double* Jp, const double* wgt, double lam, const double* dc,
                          const double* pts, double* pts_trial, double* dp, void* cuda_stream);
/* per observation model reduction -(w r J d) - 1/2 w (J d)^2 */
int acino_sba_pred_dev(acino_handle* h, int n_obs, const int32_t* cam_idx, const int32_t* pt_idx,
                       const double* res, const double* Jc, const double* Jp, const double* wgt,
                       const double* dc, const double* dp, double* pred, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* ACINO_B200_H */
