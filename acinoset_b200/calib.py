"""Drop-in camera-geometry entry points with the reference's names and signatures
(/root/reference/src/calib/calib.py), running on libacino_b200.so.

    project_points_fisheye(obj_pts, k, d, r, t) -> (n,2) float64            calib.py:132-136
    triangulate_points_fisheye(img_pts_1, img_pts_2, k1,d1,r1,t1, k2,d2,r2,t2) -> (n,3)   :121-130
    project_points / triangulate_points / create_undistort_point_function (pinhole twins)   :25-30,52-66
    get_pairwise_3d_points_from_df(points_2d_df, k_arr, d_arr, r_arr, t_arr, triangulate_func)
        -> DataFrame[frame, marker, x, y, z]                                :394-423
The SBA entry points of the same reference module live in acinoset_b200.sba and are re-exported
here under their reference names.
"""
import numpy as np

from . import fte as _fte
from .rotations import rodrigues_to_mat, rodrigues_to_vec


def project_points_fisheye(obj_pts, k, d, r, t, device=0):
    """Fisheye projection of (n,3) or (3,) world points; d may be (4,) or (4,1), t (3,1) or (3,).
    No geometry validation: a point behind the camera is projected like any other (as in the
    reference)."""
    obj_pts = np.asarray(obj_pts, dtype=np.float64).reshape((-1, 3))
    # calib.py:134: r -> cv2.Rodrigues -> rvec -> (inside cv2.fisheye.projectPoints) matrix again, which
    # projects a slightly non-orthonormal scene matrix onto SO(3); reproduced on the host (9 numbers)
    r = rodrigues_to_mat(rodrigues_to_vec(r))
    return _fte.get_handle(device).project_points(obj_pts, k, np.asarray(d).reshape(-1)[:4], r, t)


def undistort_points_fisheye(pts, k, d, device=0):
    """cv2.fisheye.undistortPoints(pts, k, d) (normalised coordinates, default criteria)."""
    pts = np.asarray(pts, dtype=np.float64)
    out = _fte.get_handle(device).undistort_points(pts.reshape(-1, 2), k, np.asarray(d).reshape(-1)[:4])
    return out.reshape(pts.shape)


def triangulate_points_fisheye(img_pts_1, img_pts_2, k1, d1, r1, t1, k2, d2, r2, t2, device=0):
    """Two-view DLT after fisheye undistortion; inputs of any shape are read as (-1,2);
    a single point returns shape (1,3) (calib.py:293-296)."""
    h = _fte.get_handle(device)
    return h.triangulate_points(np.asarray(img_pts_1, dtype=np.float64).reshape(-1, 2),
                                np.asarray(img_pts_2, dtype=np.float64).reshape(-1, 2),
                                (k1, np.asarray(d1).reshape(-1)[:4], r1, t1),
                                (k2, np.asarray(d2).reshape(-1)[:4], r2, t2))


# ---- standard (pinhole) camera model twins, calib.py:25-30,52-66 ----------------------------------
def _as_rmat(r):
    """cv2.projectPoints takes a rotation matrix or a Rodrigues vector; a matrix goes through
    cv2.Rodrigues -> rvec -> matrix inside OpenCV (projection onto SO(3)), reproduced here."""
    r = np.asarray(r, dtype=np.float64)
    if r.size == 9:
        return rodrigues_to_mat(rodrigues_to_vec(r.reshape(3, 3)))
    return rodrigues_to_mat(r.reshape(3))


def project_points(obj_pts, k, d, r, t, device=0):
    """cv2.projectPoints(obj_pts, r, t, k, d)[0].reshape(-1, 2) (calib.py:64-66); d = up to 12 OpenCV
    coefficients (plumb-bob, rational model of calib.py:18, thin prism) or None."""
    obj_pts = np.asarray(obj_pts, dtype=np.float64).reshape((-1, 3))
    return _fte.get_handle(device).project_points_pinhole(obj_pts, k, d, _as_rmat(r), t)


def create_undistort_point_function(k, d, device=0):
    """calib.py:25-30: pixels -> undistorted pixels (cv2.undistortPoints(pts, k, d, P=k))."""
    def undistort_points(pts):
        pts = np.asarray(pts, dtype=np.float64).reshape((-1, 2))
        return _fte.get_handle(device).undistort_points_pinhole(pts, k, d, to_pixels=True)
    return undistort_points


def triangulate_points(img_pts_1, img_pts_2, k1, d1, r1, t1, k2, d2, r2, t2, device=0):
    """Two-view DLT after cv2.undistortPoints (calib.py:52-61); r is used as given ([r | t], no Rodrigues)."""
    h = _fte.get_handle(device)
    return h.triangulate_points_pinhole(np.asarray(img_pts_1, dtype=np.float64).reshape(-1, 2),
                                        np.asarray(img_pts_2, dtype=np.float64).reshape(-1, 2),
                                        (k1, d1, r1, t1), (k2, d2, r2, t2))


def triangulate_pairwise_dense(uv, valid, k_arr, d_arr, r_arr, t_arr, device=0):
    """Dense-tensor TRI: uv (N,C,L,2), valid (N,C,L) -> (pos (N,L,3) NaN-filled, count (N,L))."""
    h = _fte.set_scene(k_arr, d_arr, r_arr, t_arr, device)
    return h.triangulate_pairwise(uv, valid)


def get_pairwise_3d_points_from_df(points_2d_df, k_arr, d_arr, r_arr, t_arr, triangulate_func=None, device=0):
    """Reference signature (calib.py:394-423).  ``triangulate_func`` selects the camera model like in the reference:
    ``triangulate_points_fisheye`` (or None): the adjacent pairs (i, i+1) are triangulated by one fused kernel
    (undistort + DLT + fixed-order mean); ``triangulate_points``: the standard model, one two-view launch per adjacent
    pair and the same fixed-order mean on the host; any other callable raises (no per-point host callback path).
    Prints the same "Found N pairwise points ..." lines as the reference."""
    import pandas as pd

    from .sba import PINHOLE, camera_model

    model = camera_model(None, triangulate_func)

    df = points_2d_df
    n_cameras = len(k_arr)
    if len(df) and "lab" in str(df["frame"].iloc[0]):   # calib.py:398-400 (label-file frame names)
        df = df.copy()
        df["frame"] = df["frame"].str.replace(r".*img", "", regex=True).str.replace(".png", "", regex=False)
    frames, f_idx = np.unique(df["frame"].to_numpy(), return_inverse=True)
    markers, m_idx = np.unique(df["marker"].to_numpy().astype(str), return_inverse=True)
    cams = df["camera"].to_numpy().astype(np.int64)
    N, L = len(frames), len(markers)
    uv = np.zeros((N, n_cameras, L, 2), np.float64)
    valid = np.zeros((N, n_cameras, L), np.uint8)
    inside = (cams >= 0) & (cams < n_cameras)
    uv[f_idx[inside], cams[inside], m_idx[inside], 0] = df["x"].to_numpy(dtype=np.float64)[inside]
    uv[f_idx[inside], cams[inside], m_idx[inside], 1] = df["y"].to_numpy(dtype=np.float64)[inside]
    valid[f_idx[inside], cams[inside], m_idx[inside]] = 1
    for c in range(n_cameras - 1):
        n_pair = int(np.count_nonzero(valid[:, c] & valid[:, c + 1]))
        if n_pair:
            print(f"Found {n_pair} pairwise points between camera {c} and {c + 1}")
        else:
            print(f"No pairwise points between camera {c} and {c + 1}")
    if N * L == 0:
        return pd.DataFrame(columns=["frame", "marker", "x", "y", "z"])
    if model == PINHOLE:
        pos = np.zeros((N, L, 3))
        cnt = np.zeros((N, L), np.int64)
        for c in range(n_cameras - 1):
            both = (valid[:, c] & valid[:, c + 1]).astype(bool)
            if not both.any():
                continue
            X = triangulate_points(uv[:, c][both], uv[:, c + 1][both], k_arr[c], d_arr[c], r_arr[c], t_arr[c], k_arr[c + 1],
                                   d_arr[c + 1], r_arr[c + 1], t_arr[c + 1], device=device)
            pos[both] += X
            cnt[both] += 1
        pos = np.where(cnt[..., None] > 0, pos / np.maximum(cnt, 1)[..., None], np.nan)
    else:
        pos, cnt = triangulate_pairwise_dense(uv, valid, k_arr, d_arr, r_arr, t_arr, device)
    fi, mi = np.nonzero(cnt > 0)
    return pd.DataFrame({"frame": frames[fi], "marker": markers[mi],
                         "x": pos[fi, mi, 0], "y": pos[fi, mi, 1], "z": pos[fi, mi, 2]})


_SBA_NAMES = (
    "create_bundle_adjustment_jacobian_sparsity_matrix", "prepare_calib_board_data_for_bundle_adjustment",
    "prepare_manual_points_for_bundle_adjustment", "params_to_points_only", "cost_func_points_only",
    "bundle_adjust_board_points_only", "bundle_adjust_points_only", "params_to_points_extrinsics",
    "cost_func_points_extrinsics", "bundle_adjust_board_points_and_extrinsics", "bundle_adjust_points_and_extrinsics",
)


def __getattr__(name):   # the reference keeps the SBA functions in calib.py too
    if name in _SBA_NAMES:
        from . import sba

        return getattr(sba, name)
    raise AttributeError(name)
