"""Oracle (test infrastructure): sparse-bundle-adjustment residuals, analytic Jacobian,
sparsity pattern and problem assembly, fp64 NumPy.

Follows /root/reference/src/calib/calib.py
  * create_bundle_adjustment_jacobian_sparsity_matrix   :196-207
  * prepare_calib_board_data_for_bundle_adjustment      :210-263
  * params_to_points_only / cost_func_points_only       :307-316
  * params_to_points_extrinsics / cost_func_points_extrinsics :345-359
  * bundle_adjust_points_and_extrinsics x0 layout       :373-375
Parameter vector: [rvec_0..rvec_{C-1} | t_0..t_{C-1} | X_0..X_{n-1}]; residual vector
[u0,v0,u1,v1,...].  All cameras free (gauge not fixed), like the reference.
"""
import numpy as np

from . import fisheye, triangulate


def params_to_points_extrinsics(params, n_cameras, n_points):
    r_end = n_cameras * 3
    t_end = r_end + n_cameras * 3
    r_vecs = params[:r_end].reshape(n_cameras, 3)
    r_arr = np.array([fisheye.rodrigues(r) for r in r_vecs])
    t_arr = params[r_end:t_end].reshape(n_cameras, 3)
    obj_pts = params[t_end:].reshape(n_points, 3)
    return obj_pts, r_arr, t_arr


def pack_params(r_arr, t_arr, points_3d):
    r_vecs = np.array([fisheye.rodrigues_inv(r) for r in r_arr]).ravel()
    return np.concatenate([r_vecs, np.asarray(t_arr, dtype=np.float64).ravel(),
                           np.asarray(points_3d, dtype=np.float64).ravel()])


def cost_func_points_extrinsics(params, n_cameras, n_points, point_3d_indices, camera_indices,
                                k_arr, d_arr, points_2d):
    obj_pts, r_arr, t_arr = params_to_points_extrinsics(params, n_cameras, n_points)
    out = np.empty((len(point_3d_indices), 2))
    for c in range(n_cameras):
        m = camera_indices == c
        if m.any():
            out[m] = fisheye.project(obj_pts[point_3d_indices[m]], k_arr[c], d_arr[c], r_arr[c], t_arr[c])
    return (out - points_2d).ravel()


def cost_func_points_only(params, n_points, point_3d_indices, camera_indices, k_arr, d_arr, r_arr,
                          t_arr, points_2d):
    obj_pts = params.reshape(n_points, 3)
    out = np.empty((len(point_3d_indices), 2))
    for c in range(len(k_arr)):
        m = camera_indices == c
        if m.any():
            out[m] = fisheye.project(obj_pts[point_3d_indices[m]], k_arr[c], d_arr[c], r_arr[c], t_arr[c])
    return (out - points_2d).ravel()


def jac_blocks_points_extrinsics(params, n_cameras, n_points, point_3d_indices, camera_indices,
                                 k_arr, d_arr):
    """Analytic per-observation blocks: Jr (n_obs,2,3) d/d rvec_cam, Jt (n_obs,2,3) d/d t_cam,
    Jx (n_obs,2,3) d/d X_point."""
    obj_pts, r_arr, t_arr = params_to_points_extrinsics(params, n_cameras, n_points)
    r_vecs = params[:n_cameras * 3].reshape(n_cameras, 3)
    n_obs = len(point_3d_indices)
    Jr = np.empty((n_obs, 2, 3))
    Jt = np.empty((n_obs, 2, 3))
    Jx = np.empty((n_obs, 2, 3))
    for c in range(n_cameras):
        m = camera_indices == c
        if not m.any():
            continue
        X = obj_pts[point_3d_indices[m]]
        _, Jw, Jc = fisheye.project_jac(X, k_arr[c], d_arr[c], r_arr[c], t_arr[c])
        dR = fisheye.drodrigues(r_vecs[c])  # (3,3,3): dR[i,j]/dr[k]
        dXc = np.einsum("ijk,nj->nik", dR, X)  # d(R X)/d rvec
        Jr[m] = Jc @ dXc
        Jt[m] = Jc
        Jx[m] = Jw
    return Jr, Jt, Jx


def dense_jacobian(Jr, Jt, Jx, n_cameras, n_points, point_3d_indices, camera_indices):
    n_obs = len(point_3d_indices)
    J = np.zeros((2 * n_obs, 6 * n_cameras + 3 * n_points))
    for i in range(n_obs):
        c, p = camera_indices[i], point_3d_indices[i]
        J[2 * i:2 * i + 2, 3 * c:3 * c + 3] = Jr[i]
        J[2 * i:2 * i + 2, 3 * n_cameras + 3 * c:3 * n_cameras + 3 * c + 3] = Jt[i]
        J[2 * i:2 * i + 2, 6 * n_cameras + 3 * p:6 * n_cameras + 3 * p + 3] = Jx[i]
    return J


def sparsity(n_cameras, n_params_per_camera, camera_indices, n_points, point_indices):
    """Dense 0/1 restatement of create_bundle_adjustment_jacobian_sparsity_matrix (calib.py:196-207).

    NOTE the reference's column layout here is camera-major blocks of n_params_per_camera
    (calib.py:202), which differs from the parameter layout [all rvecs | all tvecs | points]
    used by params_to_points_extrinsics (calib.py:346-351); reproduced as is.
    """
    m = camera_indices.size * 2
    n = n_cameras * n_params_per_camera + n_points * 3
    A = np.zeros((m, n), dtype=np.int8)
    i = np.arange(camera_indices.size)
    for s in range(n_params_per_camera):
        A[2 * i, camera_indices * n_params_per_camera + s] = 1
        A[2 * i + 1, camera_indices * n_params_per_camera + s] = 1
    for s in range(3):
        A[2 * i, n_cameras * n_params_per_camera + point_indices * 3 + s] = 1
        A[2 * i + 1, n_cameras * n_params_per_camera + point_indices * 3 + s] = 1
    return A


def prepare_calib_board_data(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr, r_arr, t_arr,
                             view_order=None):
    """prepare_calib_board_data_for_bundle_adjustment (calib.py:210-263).

    The reference iterates a Python ``set`` of file names (order not deterministic);
    here views are visited in ``view_order`` if given, else sorted by name.
    """
    n_cam = len(img_pts_arr)
    counts = {}
    for fnames in fnames_arr:
        for f in fnames:
            counts[f] = counts.get(f, 0) + 1
    views = [f for f, v in counts.items() if v >= 2]
    views = sorted(views) if view_order is None else [f for f in view_order if f in set(views)]
    ppi = board_shape[0] * board_shape[1]
    points_3d, point_3d_indices, points_2d, camera_indices = [], [], [], []
    counter = 0
    for fname in views:
        tri_pt, tri_cam = [], []
        for cam_idx in range(n_cam):
            if fname in fnames_arr[cam_idx]:
                f_idx = fnames_arr[cam_idx].index(fname)
                tri_pt.append(f_idx)
                tri_cam.append(cam_idx)
                points_2d.extend(np.asarray(img_pts_arr[cam_idx][f_idx]).reshape(ppi, 2))
                point_3d_indices.extend(range(counter, counter + ppi))
                camera_indices.extend([cam_idx] * ppi)
        a, b = tri_cam[0], tri_cam[1]
        est = triangulate.triangulate_points_fisheye(
            img_pts_arr[a][tri_pt[0]], img_pts_arr[b][tri_pt[1]],
            k_arr[a], d_arr[a], r_arr[a], t_arr[a], k_arr[b], d_arr[b], r_arr[b], t_arr[b])
        points_3d.extend(est)
        counter += ppi
    return (np.array(points_2d, dtype=np.float32), np.array(points_3d, dtype=np.float32),
            np.array(point_3d_indices, dtype=np.int64), np.array(camera_indices, dtype=np.int64), views)
