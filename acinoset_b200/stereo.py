"""Pairwise extrinsic calibration with the reference's names
(/root/reference/src/calib/calib.py:125-134,141-194; src/calib/app.py:84-124):

    calibrate_pair_extrinsics_fisheye(obj_pts, img_pts_1, img_pts_2, k_1, d_1, k_2, d_2, camera_resolution) -> (rms, r, t)
    calibrate_pair_extrinsics(...)  - the standard-camera-model twin (calib.py:41-49, cv2.stereoCalibrate)
    calibrate_pairwise_extrinsics(calib_func, img_pts_arr, fnames_arr, k_arr, d_arr, camera_resolution, board_shape,
                                  board_edge_len) -> (r_arr, t_arr)
    calibrate_fisheye_extrinsics_pairwise(camera_fpaths, points_fpaths, out_fpath)

The reference hands the pair problem to cv2.fisheye.stereoCalibrate (CALIB_FIX_INTRINSIC, 100 iterations, eps 1e-5).
Here every arithmetic step runs in CUDA kernels (csrc/stereo.cu): per-view board poses from the planar homography,
then Levenberg-Marquardt on the joint reprojection error of both cameras with the per-view blocks eliminated; this
module only drives the accept / reject loop and chains the pairs (calib.py:141-194).  rms follows OpenCV's definition:
sqrt(sum of squared pixel errors / (2 n_views n_points)).
"""
import ctypes

import numpy as np

from . import _lib
from . import fte as _fte
from .rotations import rodrigues_to_mat, rodrigues_to_vec


class _GpuBackend:
    """acino_stereo_* of libacino_b200.so (no CPU fallback)."""

    def __init__(self, device=0):
        self.h = _fte.get_handle(device)

    def set(self, obj, img1, img2, K1, D1, K2, D2, pinhole=False):
        self.V, self.M = img1.shape[0], img1.shape[1]
        p = _lib._np_ptr
        if pinhole:
            self.h._check(_lib.lib.acino_stereo_set_pinhole(self.h._h, self.V, self.M, p(obj), p(img1), p(img2), p(K1), p(D1), D1.size,
                                                            p(K2), p(D2), D2.size), "acino_stereo_set_pinhole")
        else:
            self.h._check(_lib.lib.acino_stereo_set(self.h._h, self.V, self.M, p(obj), p(img1), p(img2), p(K1), p(D1), p(K2), p(D2)),
                          "acino_stereo_set")

    def init(self):
        poses = np.empty((self.V, 2, 12))
        cost = np.empty((self.V, 2))
        self.h._check(_lib.lib.acino_stereo_init(self.h._h, _lib._np_ptr(poses), _lib._np_ptr(cost)), "acino_stereo_init")
        return poses, cost

    def step(self, rel, poses, lam):
        rel_t, poses_t = np.empty(12), np.empty((self.V, 12))
        cost, info = ctypes.c_double(), ctypes.c_int32()
        self.h._check(_lib.lib.acino_stereo_step(self.h._h, _lib._np_ptr(rel), _lib._np_ptr(poses), float(lam), _lib._np_ptr(rel_t),
                                                 _lib._np_ptr(poses_t), ctypes.byref(cost), ctypes.byref(info)), "acino_stereo_step")
        return rel_t, poses_t, cost.value, info.value


def _median_relative_pose(poses):
    """poses (V, 2, 12) board poses in camera 1 / camera 2 -> initial (R, T) of camera 2 w.r.t. camera 1: the
    component-wise median of the per-view relative rotation vectors and translations (OpenCV's choice too)."""
    rv, tv = [], []
    for p1, p2 in poses:
        R1, t1, R2, t2 = p1[:9].reshape(3, 3), p1[9:], p2[:9].reshape(3, 3), p2[9:]
        Rr = R2 @ R1.T
        rv.append(rodrigues_to_vec(Rr).ravel())
        tv.append(t2 - Rr @ t1)
    return rodrigues_to_mat(np.median(np.array(rv), axis=0)), np.median(np.array(tv), axis=0)


def solve_pair(obj_pts, img_pts_1, img_pts_2, k_1, d_1, k_2, d_2, max_iter=100, eps=1e-10, device=0,
               return_info=False, pinhole=False):
    """-> (rms, R (3,3), T (3,1)) [, info].  img_pts_* (V, M, 2) pixels of the same V views, obj_pts (M, 3).
    pinhole=True: d_* are OpenCV's standard-model coefficients (up to 12) instead of the 4 fisheye ones.
    There is no CPU path: a missing GPU raises AcinoError."""
    obj = np.ascontiguousarray(obj_pts, dtype=np.float64).reshape(-1, 3)
    M = obj.shape[0]
    img1 = np.ascontiguousarray(img_pts_1, dtype=np.float64).reshape(-1, M, 2)
    img2 = np.ascontiguousarray(img_pts_2, dtype=np.float64).reshape(-1, M, 2)
    if img1.shape != img2.shape or img1.shape[0] < 1:
        raise ValueError("img_pts_1 and img_pts_2 must hold the same (>= 1) views of the same board")
    V = img1.shape[0]
    K1, K2 = (np.ascontiguousarray(k, dtype=np.float64).reshape(3, 3) for k in (k_1, k_2))
    nd = 14 if pinhole else 4
    D1, D2 = (np.ascontiguousarray(np.zeros(0) if d is None else np.asarray(d, dtype=np.float64).reshape(-1)[:nd]) for d in (d_1, d_2))
    be = _GpuBackend(device)
    be.set(obj, img1, img2, K1, D1, K2, D2, pinhole=pinhole)
    poses2, cost0 = be.init()
    if np.any(cost0 < 0):
        raise _lib.AcinoError(f"board pose initialisation failed for views {np.nonzero((cost0 < 0).any(axis=1))[0].tolist()}")
    Rr, Tr = _median_relative_pose(poses2)
    rel = np.concatenate([Rr.ravel(), Tr.ravel()])
    poses = np.ascontiguousarray(poses2[:, 0])
    _, _, F, _ = be.step(rel, poses, -1.0)
    lam, it, n_eval = 1e-3, 0, 1
    for it in range(max_iter):
        accepted = False
        for _ in range(12):
            rel_t, poses_t, Ft, info = be.step(rel, poses, lam)
            n_eval += 1
            if info == 0 and Ft < F:
                dF = (F - Ft) / max(F, 1e-300)
                rel, poses, F = rel_t, poses_t, Ft
                lam = max(lam / 10, 1e-12)
                accepted = True
                break
            lam *= 10
        if not accepted or dF < eps:
            break
    rms = float(np.sqrt(F / (2 * V * M)))
    out = (rms, rel[:9].reshape(3, 3).copy(), rel[9:].reshape(3, 1).copy())
    if return_info:
        return out + (dict(iters=it + 1, evals=n_eval, cost=F, init_cost=cost0, poses=poses),)
    return out


def calibrate_pair_extrinsics_fisheye(obj_pts, img_pts_1, img_pts_2, k_1, d_1, k_2, d_2, camera_resolution=None, device=0):
    """calib.py:125-134.  img_pts_* (n_views, rows, cols, 2) or (n_views, M, 2)."""
    n = np.asarray(img_pts_1).shape[0]
    return solve_pair(obj_pts, np.asarray(img_pts_1).reshape(n, -1, 2), np.asarray(img_pts_2).reshape(n, -1, 2), k_1, d_1, k_2,
                      d_2, device=device)


def calibrate_pair_extrinsics(obj_pts, img_pts_1, img_pts_2, k_1, d_1, k_2, d_2, camera_resolution=None, device=0,
                              rational_model=False):
    """calib.py:41-49 (cv2.stereoCalibrate, flags = CALIB_FIX_INTRINSIC only): the standard-camera-model twin.

    Reference behaviour kept by default: without CALIB_RATIONAL_MODEL / CALIB_THIN_PRISM_MODEL in the flags OpenCV drops
    k4..k6 and s1..s4 of the distortion vectors it is given, so the reference's pair calibration runs on the 5-coefficient
    model even though its intrinsics were calibrated with 8 (calib.py:18) - verified against cv2 4.13.0 to 1e-8
    (tests/golden/stereo.npz).  ``rational_model=True`` uses every coefficient."""
    n = np.asarray(img_pts_1).shape[0]

    def coeffs(d):
        d = np.zeros(0) if d is None else np.asarray(d, dtype=np.float64).reshape(-1)
        return d if rational_model else d[:5]

    return solve_pair(obj_pts, np.asarray(img_pts_1).reshape(n, -1, 2), np.asarray(img_pts_2).reshape(n, -1, 2), k_1, coeffs(d_1),
                      k_2, coeffs(d_2), device=device, pinhole=True)


def calibrate_pairwise_extrinsics(calib_func, img_pts_arr, fnames_arr, k_arr, d_arr, camera_resolution, board_shape,
                                  board_edge_len):
    """calib.py:141-194: camera 1 at R1 = [[1,0,0],[0,0,-1],[0,1,0]], T1 = 0 (world z up), every next camera placed by
    the relative pose of the pair (i, i+1) estimated from the views both saw."""
    from .utils import create_board_object_pts

    n_cam = len(img_pts_arr)
    R1 = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float64)
    T1 = np.zeros((3, 1))
    r_arr, t_arr = [R1], [T1]
    obj_pts = create_board_object_pts(board_shape, board_edge_len)
    for i in range(n_cam - 1):
        names_2 = {f: b for b, f in enumerate(fnames_arr[i + 1])}
        pairs = [(a, names_2[f]) for a, f in enumerate(fnames_arr[i]) if f in names_2]
        assert pairs, f"No corresponding points between img_pts at index {i} and {i + 1}"
        pts_1 = np.array([img_pts_arr[i][a] for a, _ in pairs], dtype=np.float32)
        pts_2 = np.array([img_pts_arr[i + 1][b] for _, b in pairs], dtype=np.float32)
        rms, r, t = calib_func(obj_pts, pts_1, pts_2, k_arr[i], d_arr[i], k_arr[i + 1], d_arr[i + 1], camera_resolution)
        print(f"{i + 1} & {i + 2}\t\t{len(pairs)}\t\t{rms:.5f} pixels")
        R1, T1 = r @ R1, r @ T1 + np.asarray(t).reshape(3, 1)
        r_arr.append(R1)
        t_arr.append(T1)
    return r_arr, t_arr


def calibrate_standard_extrinsics_pairwise(camera_fpaths, points_fpaths, out_fpath, device=0):
    """app.py:119-120"""
    return calibrate_fisheye_extrinsics_pairwise(camera_fpaths, points_fpaths, out_fpath, device=device,
                                                 _pair_func=calibrate_pair_extrinsics)


def calibrate_fisheye_extrinsics_pairwise(camera_fpaths, points_fpaths, out_fpath, device=0, _pair_func=None):
    """app.py:84-124: camera files + point files -> scene file with the chained pairwise extrinsics."""
    from . import utils

    k_arr, d_arr, cam_res = [], [], None
    for c in camera_fpaths:
        k, d, res = utils.load_camera(c)
        k_arr.append(k)
        d_arr.append(d)
        assert cam_res is None or tuple(cam_res) == tuple(res)
        cam_res = res
    img_pts_arr, fnames_arr, board_shape, board_edge_len = [], [], None, None
    for p in points_fpaths:
        points, fnames, bs, bel, _ = utils.load_points(p)
        img_pts_arr.append(points)
        fnames_arr.append(fnames)
        assert board_shape is None or tuple(board_shape) == tuple(bs)
        board_shape, board_edge_len = bs, bel
    print("camera pair\tcommon frames\tRMS reprojection error")
    pair = calibrate_pair_extrinsics_fisheye if _pair_func is None else _pair_func
    r_arr, t_arr = calibrate_pairwise_extrinsics(lambda *a: pair(*a, device=device), img_pts_arr,
                                                 fnames_arr, k_arr, d_arr, cam_res, board_shape, board_edge_len)
    utils.save_scene(out_fpath, np.array(k_arr), np.array(d_arr), np.array(r_arr), np.array(t_arr), cam_res)
    return r_arr, t_arr
