"""The reference's pipeline script (/root/reference/src/all_optimizations.py) by name: ``tri``, ``sba``, ``ekf``,
``fte`` over one data directory, and the scalar projection helpers ``pt3d_to_2d`` / ``pt3d_to_x2d`` /
``pt3d_to_y2d`` (:193-217).  Every stage runs on libacino_b200.so; only file handling and the (sequential) Kalman
recursion stay on the host.  Plotting, video labelling and logging of the reference are out of scope
 and are not reproduced.

Directory layout the reference expects: ``DATA_DIR/dlc/*.h5`` (or the sibling ``.csv``), a ``*_cam_scene*_sba.json``
in an ``extrinsic_calib`` folder at or above DATA_DIR, outputs in ``DATA_DIR/{tri,sba,ekf,fte}/``.
"""
import os
import pickle
from glob import glob
from time import time

import numpy as np

from . import calib, utils
from . import ekf as _ekf
from . import fte as _fte
from .fte import MARKERS, fte  # noqa: F401  (all_optimizations.py:22)


def pt3d_to_2d(x, y, z, K, D, R, t, device=0):
    """all_optimizations.py:193-209: fisheye projection of one world point (or arrays of them) -> (u, v).
    R is used as given (no Rodrigues round trip, unlike project_points_fisheye)."""
    X = np.stack(np.broadcast_arrays(np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(z, np.float64)), -1)
    uv = _fte.get_handle(device).project_points(X.reshape(-1, 3), K, np.asarray(D).reshape(-1)[:4], R, t)
    uv = uv.reshape(X.shape[:-1] + (2,))
    return (float(uv[0]), float(uv[1])) if uv.ndim == 1 else (uv[..., 0], uv[..., 1])


def pt3d_to_x2d(x, y, z, K, D, R, t, device=0):
    """:211-213"""
    return pt3d_to_2d(x, y, z, K, D, R, t, device)[0]


def pt3d_to_y2d(x, y, z, K, D, R, t, device=0):
    """:215-217"""
    return pt3d_to_2d(x, y, z, K, D, R, t, device)[1]


def _dlc_paths(dlc_dir):
    return sorted(glob(os.path.join(dlc_dir, "*.h5"))) or sorted(glob(os.path.join(dlc_dir, "*.csv")))


def _positions_from_df(points_3d_df, markers, start_frame, n_frames):
    """:885-890 / :929-934: long [frame, marker, x, y, z] -> dense (N, L, 3), NaN where missing."""
    positions = np.full((n_frames, len(markers), 3), np.nan)
    mi = {m: i for i, m in enumerate(markers)}
    fr = points_3d_df["frame"].to_numpy().astype(np.int64) - start_frame
    mk = np.array([mi.get(m, -1) for m in points_3d_df["marker"].to_numpy()], dtype=np.int64)
    ok = (fr >= 0) & (fr < n_frames) & (mk >= 0)
    positions[fr[ok], mk[ok]] = points_3d_df[["x", "y", "z"]].to_numpy(dtype=np.float64)[ok]
    return positions


def _save_positions(positions, out_dir, name, scene_fpath, start_frame, device):
    """app.save_tri / save_sba (:893,936): ``{name}.pickle`` {positions, start_frame} + the 2-D reprojection files."""
    os.makedirs(out_dir, exist_ok=True)
    out_fpath = os.path.join(out_dir, f"{name}.pickle")
    with open(out_fpath, "wb") as f:
        pickle.dump(dict(positions=positions, start_frame=start_frame), f)
    print(f"Saved {out_fpath}")
    _fte.save_3d_cheetah_as_2d(positions, out_dir, scene_fpath, MARKERS, None, start_frame, out_fname=name,
                               device=device)
    return out_fpath


def _load(DATA_DIR, start_frame, end_frame):
    assert os.path.exists(DATA_DIR)
    dlc_dir = os.path.join(DATA_DIR, "dlc")
    assert os.path.exists(dlc_dir)
    k_arr, d_arr, r_arr, t_arr, cam_res, n_cams, scene_fpath = utils.find_scene_file(DATA_DIR, verbose=False)
    paths = _dlc_paths(dlc_dir)
    assert n_cams == len(paths), f"# of dlc files != # of cams in {scene_fpath}"
    df = utils.load_dlc_points_as_df(paths, verbose=False)
    tot_frames = int(df["frame"].max()) + 1
    assert end_frame <= tot_frames, f"end_frame must be less than or equal to {tot_frames}"
    end_frame = tot_frames if end_frame == -1 else end_frame
    start_frame -= 1          # 0 based indexing (:59,607)
    assert start_frame >= 0
    return df, (k_arr, d_arr.reshape(-1, 4), r_arr, t_arr), cam_res, scene_fpath, start_frame, end_frame


def tri(DATA_DIR, start_frame, end_frame, dlc_thresh, device=0):
    """:906-936: pairwise DLT triangulation of every (frame, marker) -> ``tri/tri.pickle``."""
    df, (k, d, r, t), _, scene_fpath, start_frame, end_frame = _load(DATA_DIR, start_frame, end_frame)
    N = end_frame - start_frame
    df = df[df["frame"].between(start_frame, end_frame - 1)]
    df = df[df["likelihood"] > dlc_thresh]
    assert len(k) == df["camera"].nunique()
    points_3d_df = calib.get_pairwise_3d_points_from_df(df, k, d, r, t, calib.triangulate_points_fisheye, device=device)
    positions = _positions_from_df(points_3d_df, MARKERS, start_frame, N)
    _save_positions(positions, os.path.join(DATA_DIR, "tri"), "tri", scene_fpath, start_frame, device)
    return positions


def sba(DATA_DIR, start_frame, end_frame, dlc_thresh, device=0):
    """:868-893 (called with these four arguments at :969): point-only bundle adjustment of the triangulated
    markers against every camera that saw them -> ``sba/sba.pickle``."""
    from . import sba as _sba

    df, (k, d, r, t), _, scene_fpath, start_frame, end_frame = _load(DATA_DIR, start_frame, end_frame)
    N = end_frame - start_frame
    df = df[df["frame"].between(start_frame, end_frame - 1)]
    df = df[df["likelihood"] > dlc_thresh].reset_index(drop=True)
    t0 = time()
    points_3d_df, residuals = _sba.sba_points_fisheye(scene_fpath, df, device=device)
    print("\nOptimization took {0:.2f} seconds\n".format(time() - t0))
    positions = _positions_from_df(points_3d_df, MARKERS, start_frame, N)
    _save_positions(positions, os.path.join(DATA_DIR, "sba"), "sba", scene_fpath, start_frame, device)
    return positions, residuals


def ekf(DATA_DIR, start_frame, end_frame, dlc_thresh, fps=None, device=0):
    """:569-866: EKF + RTS smoother -> ``ekf/ekf.pickle`` {positions, x, dx, ddx, smoothed_x, smoothed_dx,
    smoothed_ddx, start_frame}.  The measurement Jacobian of every frame is one fte_jac launch."""
    from scipy.stats import linregress

    t0 = time()
    df, (k, d, r, t), cam_res, scene_fpath, start_frame, end_frame = _load(DATA_DIR, start_frame, end_frame)
    if fps is None:
        import cv2

        vids = sorted(glob(os.path.join(DATA_DIR, "cam[1-9].mp4")))
        assert vids, "fps not given and no cam[1-9].mp4 next to the data"
        cap = cv2.VideoCapture(vids[0])
        fps = cap.get(cv2.CAP_PROP_FPS)
        cap.release()
    n_frames = end_frame - start_frame
    sT = 1.0 / fps
    _fte.set_scene(k, d, r, t, device)
    points_3d_df = calib.get_pairwise_3d_points_from_df(df[df["likelihood"] > dlc_thresh], k, d, r, t,
                                                        calib.triangulate_points_fisheye, device=device)
    meas, lik = utils.dlc_df_to_dense(df, len(k), MARKERS, start_frame, n_frames)
    pixels_arr = meas.reshape(n_frames, -1).astype(np.float64)       # [camera][marker][x, y] (:668-670)
    likelihood_arr = lik.reshape(n_frames, -1).astype(np.float64)
    # initial states (:691-706): head x, y, yaw and planar velocity from the triangulated nose
    idx = _ekf.get_pose_params()
    n = _ekf.N_POSE
    points_3d_df = points_3d_df[points_3d_df["frame"].between(start_frame, end_frame - 1)]
    nose = points_3d_df[points_3d_df["marker"] == "nose"][["frame", "x", "y", "z"]].to_numpy(dtype=np.float64)
    xs, xi, *_ = linregress(nose[:, 0], nose[:, 1])
    ys, yi, *_ = linregress(nose[:, 0], nose[:, 2])
    states = np.zeros(3 * n)
    states[[idx["x_0"], idx["y_0"], idx["psi_0"]]] = [start_frame * xs + xi, start_frame * ys + yi, np.arctan2(ys, xs)]
    states[[n + idx["x_0"], n + idx["y_0"]]] = [xs / sT, ys / sT]
    print("\nInitialization took {0:.2f} seconds\n".format(time() - t0))
    t1 = time()
    out = _ekf.ekf_filter(pixels_arr, likelihood_arr, states, fps, dlc_thresh, cam_res[0], device=device)
    print("\nOptimization took {0:.2f} seconds\n".format(time() - t1))
    out["positions"] = _ekf.get_3d_marker_coords(out["smoothed_x"], device)          # :849-853
    out["start_frame"] = start_frame
    out_dir = os.path.join(DATA_DIR, "ekf")
    os.makedirs(out_dir, exist_ok=True)
    out_fpath = os.path.join(out_dir, "ekf.pickle")
    with open(out_fpath, "wb") as f:
        pickle.dump(out, f)
    print(f"Saved {out_fpath}")
    _fte.save_3d_cheetah_as_2d(out["positions"], out_dir, scene_fpath, MARKERS, None, start_frame, out_fname="ekf", device=device)
    return out


def main(argv=None):
    """:941-975 without the DLC video labelling stage."""
    from argparse import ArgumentParser

    parser = ArgumentParser(description="All Optimizations")
    parser.add_argument("--data_dir", type=str, required=True)
    parser.add_argument("--start_frame", type=int, default=1)
    parser.add_argument("--end_frame", type=int, default=-1)
    parser.add_argument("--dlc_thresh", type=float, default=0.8)
    parser.add_argument("--fps", type=float, default=None)
    args = parser.parse_args(argv)
    print("========== Triangulation ==========\n")
    tri(args.data_dir, args.start_frame, args.end_frame, args.dlc_thresh)
    print("========== SBA ==========\n")
    sba(args.data_dir, args.start_frame, args.end_frame, args.dlc_thresh)
    print("========== EKF ==========\n")
    ekf(args.data_dir, args.start_frame, args.end_frame, args.dlc_thresh, fps=args.fps)
    print("========== FTE ==========\n")
    fte(args.data_dir, args.start_frame, args.end_frame, args.dlc_thresh, fps=args.fps)


if __name__ == "__main__":
    main()
