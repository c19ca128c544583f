"""Scene / camera / points loaders with the reference's names (/root/reference/src/calib/utils.py).

Reads both JSON key dialects: the reference code writes ``board_edge_len`` / ``created_timestamp``
(utils.py:23-27) while every shipped data file uses ``board_square_len`` / ``timestamp``
(SURVEY.md appendix A).
"""
import json
import os
from datetime import datetime
from glob import glob

import numpy as np


def create_board_object_pts(board_shape, square_edge_length):
    """utils.py:10-13"""
    object_pts = np.zeros((board_shape[0] * board_shape[1], 3), np.float32)
    object_pts[:, :2] = np.mgrid[0:board_shape[0], 0:board_shape[1]].T.reshape(-1, 2) * square_edge_length
    return object_pts


def save_points(out_fpath, img_points, img_fnames, board_shape, board_edge_len, camera_resolution):
    """utils.py:16-30"""
    if isinstance(img_points, np.ndarray):
        img_points = img_points.tolist()
    data = {"created_timestamp": str(datetime.now()), "board_shape": list(board_shape),
            "board_edge_len": board_edge_len, "camera_resolution": list(camera_resolution),
            "points": dict(zip(img_fnames, img_points))}
    with open(out_fpath, "w") as f:
        json.dump(data, f)


def load_points(fpath):
    """utils.py:33-41 -> (points (n,rows*cols or rows,cols,2) f32, fnames, board_shape, board_edge_len, resolution)"""
    with open(fpath) as f:
        data = json.load(f)
    fnames = list(data["points"].keys())
    points = np.array(list(data["points"].values()), dtype=np.float32)
    board_shape = tuple(data["board_shape"])
    board_edge_len = data["board_edge_len"] if "board_edge_len" in data else data["board_square_len"]
    return points, fnames, board_shape, board_edge_len, tuple(data["camera_resolution"])


def save_camera(out_fpath, camera_resolution, k, d):
    """utils.py:44-53"""
    data = {"created_timestamp": str(datetime.now()), "camera_resolution": list(camera_resolution),
            "k": np.asarray(k).tolist(), "d": np.asarray(d).tolist()}
    with open(out_fpath, "w") as f:
        json.dump(data, f)


def load_camera(fpath):
    """utils.py:56-62"""
    with open(fpath) as f:
        data = json.load(f)
    return (np.array(data["k"], dtype=np.float64), np.array(data["d"], dtype=np.float64),
            tuple(data["camera_resolution"]))


def save_scene(out_fpath, k_arr, d_arr, r_arr, t_arr, camera_resolution):
    """utils.py:65-81"""
    cameras = [{"k": np.asarray(k).tolist(), "d": np.asarray(d).tolist(), "r": np.asarray(r).tolist(),
                "t": np.asarray(t).tolist()} for k, d, r, t in zip(k_arr, d_arr, r_arr, t_arr)]
    data = {"created_timestamp": str(datetime.now()), "camera_resolution": list(camera_resolution), "cameras": cameras}
    with open(out_fpath, "w") as f:
        json.dump(data, f)


def load_scene(fpath):
    """utils.py:84-101 -> (k_arr (C,3,3), d_arr (C,4,1), r_arr (C,3,3), t_arr (C,3,1), camera_resolution)"""
    with open(fpath) as f:
        data = json.load(f)
    cams = data["cameras"]
    k_arr = np.array([c["k"] for c in cams], dtype=np.float64)
    d_arr = np.array([c["d"] for c in cams], dtype=np.float64)
    r_arr = np.array([c["r"] for c in cams], dtype=np.float64)
    t_arr = np.array([c["t"] for c in cams], dtype=np.float64)
    return k_arr, d_arr, r_arr, t_arr, tuple(data["camera_resolution"])


def find_scene_file(dir_path, scene_fname=None, verbose=True):
    """Newer-API loader used at all_optimizations.py:220,601,911 (lib.utils, missing from the
    snapshot; semantics from its call sites): walk up from ``dir_path`` until a
    ``*_cam_scene*_sba.json`` (or ``scene_fname``) is found in an ``extrinsic_calib`` folder.
    -> (k_arr, d_arr, r_arr, t_arr, cam_res, n_cams, scene_fpath)"""
    d = os.path.abspath(dir_path)
    while True:
        for sub in ("extrinsic_calib", "."):
            pat = os.path.normpath(os.path.join(d, sub, scene_fname or "*_cam_scene*_sba.json"))
            hits = sorted(glob(pat))
            if hits:
                k, dd, r, t, res = load_scene(hits[-1])
                if verbose:
                    print(f"Loaded extrinsics from {hits[-1]}")
                return k, dd, r, t, res, len(k), hits[-1]
        parent = os.path.dirname(d)
        if parent == d:
            raise FileNotFoundError(f"no scene file found at or above {dir_path}")
        d = parent


def _dlc_wide_to_long(dlc_df, camera):
    """One DLC table (columns MultiIndex scorer/bodyparts/coords) -> long rows for one camera."""
    import pandas as pd

    df = dlc_df.droplevel(0, axis=1) if dlc_df.columns.nlevels == 3 else dlc_df
    bodyparts = list(dict.fromkeys(df.columns.get_level_values(0)))
    n = len(df)
    frames = np.asarray(df.index)
    out = []
    for bp in bodyparts:
        out.append(pd.DataFrame({"frame": frames, "camera": camera, "marker": bp,
                                 "x": df[(bp, "x")].to_numpy(), "y": df[(bp, "y")].to_numpy(),
                                 "likelihood": df[(bp, "likelihood")].to_numpy()}))
    res = pd.concat(out, ignore_index=True)
    return res.sort_values(["frame", "marker"], kind="stable").reset_index(drop=True) if n else res


def load_dlc_points_as_df(dlc_df_fpaths, verbose=True):
    """create_dlc_points_2d_file (utils.py:105-120) / lib.utils.load_dlc_points_as_df: DeepLabCut
    per-camera tables -> long-form DataFrame [frame, camera, marker, x, y, likelihood], camera =
    0-based index in the given (sorted) list.  Reads ``.h5`` (needs pytables) or the sibling
    ``.csv`` (3-row header scorer/bodyparts/coords)."""
    import pandas as pd

    dfs = []
    for i, path in enumerate(dlc_df_fpaths):
        if str(path).endswith(".csv"):
            tab = pd.read_csv(path, header=[0, 1, 2], index_col=0)
        else:
            try:
                tab = pd.read_hdf(path)
            except ImportError:
                tab = pd.read_csv(os.path.splitext(path)[0] + ".csv", header=[0, 1, 2], index_col=0)
        dfs.append(_dlc_wide_to_long(tab, i))
    df = pd.concat(dfs, ignore_index=True) if dfs else pd.DataFrame(
        columns=["frame", "camera", "marker", "x", "y", "likelihood"])
    if verbose:
        print(f"DLC points dataframe:\n{df}")
    return df[["frame", "camera", "marker", "x", "y", "likelihood"]]


create_dlc_points_2d_file = load_dlc_points_as_df


def dlc_df_to_dense(points_2d_df, n_cams, markers, start_frame, n_frames):
    """Long-form DataFrame -> dense tensors meas (N,C,L,2) f32, likelihood (N,C,L) f32 in ONE pass
    (replaces the three boolean masks per scalar of get_meas_from_df / get_likelihood_from_df,
    all_optimizations.py:226-239).  Missing rows get likelihood 0."""
    mi = {m: i for i, m in enumerate(markers)}
    L = len(markers)
    meas = np.zeros((n_frames, n_cams, L, 2), np.float32)
    lik = np.zeros((n_frames, n_cams, L), np.float32)
    fr = points_2d_df["frame"].to_numpy().astype(np.int64) - start_frame
    cam = points_2d_df["camera"].to_numpy().astype(np.int64)
    mk = np.array([mi.get(m, -1) for m in points_2d_df["marker"].to_numpy()], dtype=np.int64)
    ok = (fr >= 0) & (fr < n_frames) & (cam >= 0) & (cam < n_cams) & (mk >= 0)
    meas[fr[ok], cam[ok], mk[ok], 0] = points_2d_df["x"].to_numpy(dtype=np.float32)[ok]
    meas[fr[ok], cam[ok], mk[ok], 1] = points_2d_df["y"].to_numpy(dtype=np.float32)[ok]
    lik[fr[ok], cam[ok], mk[ok]] = points_2d_df["likelihood"].to_numpy(dtype=np.float32)[ok]
    return meas, lik
