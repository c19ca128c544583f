"""Full Trajectory Estimation (FTE) on the GPU - host side.

Mirrors the FTE section of the reference (src/all_optimizations.py:22-566): the same
measurement model (pt3d_to_2d :193-209), weights (:302-308), redescending loss (:25-27,497),
dynamics/smoothness term (:369-391, :245-252) and pose bounds (:403-483); the residual /
Jacobian evaluation and the solve run in libacino_b200.so.
"""
import numpy as np

from . import _lib

N_ACTIVE = _lib.N_ACTIVE
N_MARKERS = _lib.N_MARKERS
N_UPPER = _lib.N_UPPER

# marker order = rows of `positions` (all_optimizations.py:170-178)
MARKERS = [
    "l_eye", "r_eye", "nose", "neck_base", "spine", "tail_base", "tail1", "tail2",
    "l_shoulder", "l_front_knee", "l_front_ankle", "r_shoulder", "r_front_knee",
    "r_front_ankle", "l_hip", "l_back_knee", "l_back_ankle", "r_hip", "r_back_knee",
    "r_back_ankle",
]
# indices of the 25 active slots in the reference's 45-vector [x,y,z,*phi(14),*theta(14),*psi(14)]
ACTIVE_IDX = np.array([0, 1, 2, 3, 4, 6] + list(range(17, 31)) + [31, 32, 34, 35, 36])
from .config import CHEETAH  # noqa: E402

MEAS_SIGMA_R = CHEETAH.meas_sigma_px  # all_optimizations.py:243

_handles = {}


def get_handle(device=0):
    """One cached library handle per GPU."""
    h = _handles.get(device)
    if h is None:
        h = _lib.Handle(device)
        _handles[device] = h
    return h


def set_scene(k_arr, d_arr, r_arr, t_arr, device=0):
    """Install the camera table (the arrays utils.load_scene / find_scene_file return)."""
    h = get_handle(device)
    h.set_cameras(k_arr, np.asarray(d_arr).reshape(-1, 4), r_arr, np.asarray(t_arr).reshape(-1, 3))
    return h


def meas_weights(likelihood, dlc_thresh, sigma=MEAS_SIGMA_R):
    """init_meas_weights (all_optimizations.py:302-308): 1/R if likelihood > thresh else 0."""
    return np.where(np.asarray(likelihood) > dlc_thresh, np.float32(1.0 / sigma), np.float32(0.0)).astype(np.float32)


def fte_eval(x, meas, w, device=0, want_H=True):
    """Residual + Jacobian evaluation of N frames through the C ABI with HOST buffers.

    x (N,25), meas (N,C,20,2), w (N,C,20) -> cost (N,), g (N,25), H (N,325) packed upper."""
    return get_handle(device).fte_eval(x, meas, w, want_H=want_H)


def pose_to_3d(x, device=0):
    """pose_to_3d (all_optimizations.py:186) for a batch of active states: (N,25) -> (N,20,3)."""
    x = np.asarray(x, dtype=np.float32)
    single = x.ndim == 1
    pos, _ = get_handle(device).fk_project(np.atleast_2d(x), want_pos=True, want_uv=False)
    return pos[0] if single else pos


def unpack_upper(Hu):
    Hu = np.asarray(Hu)
    iu = np.triu_indices(N_ACTIVE)
    H = np.zeros(Hu.shape[:-1] + (N_ACTIVE, N_ACTIVE), dtype=Hu.dtype)
    H[..., iu[0], iu[1]] = Hu
    H[..., iu[1], iu[0]] = Hu
    return H


# ---- full trajectory estimation: reference entry point -------------------------------------------
def derived_velocities(x, Ts):
    """dx, ddx of the output pickle from the optimised x (backward-Euler collocation,
    all_optimizations.py:369-383; the free leading values follow SURVEY.md appendix B6)."""
    x = np.asarray(x, dtype=np.float64)
    N = x.shape[0]
    dx = np.zeros_like(x)
    ddx = np.zeros_like(x)
    if N > 1:
        dx[1:] = (x[1:] - x[:-1]) / Ts
    if N > 2:
        ddx[2:] = (dx[2:] - dx[1:-1]) / Ts
        ddx[1] = ddx[2]
        ddx[0] = ddx[1]
    if N > 1:
        dx[0] = dx[1] - Ts * ddx[1]
    return dx, ddx


def initial_guess(points_3d_df, start_frame, end_frame, N):
    """Reference initialisation (all_optimizations.py:269-277,333-337): linear regression of the
    triangulated nose over frames -> x,y,z lines; psi0 = atan2(y_slope, x_slope); all else 0."""
    from scipy.stats import linregress

    nose = points_3d_df[points_3d_df["marker"] == "nose"][["frame", "x", "y", "z"]].values.astype(np.float64)
    if len(nose) < 2:
        raise ValueError("need at least two triangulated nose points to initialise the trajectory")
    xs, xi, *_ = linregress(nose[:, 0], nose[:, 1])
    ys, yi, *_ = linregress(nose[:, 0], nose[:, 2])
    zs, zi, *_ = linregress(nose[:, 0], nose[:, 3])
    frame_est = np.arange(end_frame)
    x0 = np.zeros((N, N_ACTIVE))
    x0[:, 0] = (frame_est * xs + xi)[start_frame:start_frame + N]
    x0[:, 1] = (frame_est * ys + yi)[start_frame:start_frame + N]
    x0[:, 2] = (frame_est * zs + zi)[start_frame:start_frame + N]
    x0[:, 20] = np.arctan2(ys, xs)        # psi0 (slot 31 of the 45-vector, all_optimizations.py:337)
    return x0


def fte_solve(points_2d_df, k_arr, d_arr, r_arr, t_arr, start_frame, end_frame, dlc_thresh, fps, device=0,
              markers=None, x0=None, verbose=False, **lm_kwargs):
    """The compute of ``fte()`` (all_optimizations.py:22-566) without the file I/O.

    points_2d_df: long-form DataFrame [frame, camera, marker, x, y, likelihood]; frames
    [start_frame, end_frame) (0-based).  Returns dict(positions (N,20,3), x, dx, ddx (N,25),
    start_frame, info) - the schema of the reference's fte.pickle (:540-559)."""
    from . import calib, lm, utils

    markers = MARKERS if markers is None else list(markers)
    N = int(end_frame - start_frame)
    Ts = 1.0 / fps
    d4 = np.asarray(d_arr, dtype=np.float64).reshape(-1, 4)
    h = set_scene(k_arr, d4, r_arr, t_arr, device)
    if x0 is None:
        pts3d = calib.get_pairwise_3d_points_from_df(points_2d_df[points_2d_df["likelihood"] > dlc_thresh], k_arr, d4,
                                                     r_arr, t_arr, calib.triangulate_points_fisheye, device=device)
        x0 = initial_guess(pts3d, start_frame, end_frame, N)
    meas, lik = utils.dlc_df_to_dense(points_2d_df, len(k_arr), markers, start_frame, N)
    w = meas_weights(lik, dlc_thresh)
    solver = lm.FTESolver(h, meas, w, Ts)
    x, info = solver.solve(x0, verbose=verbose, **lm_kwargs)
    positions = pose_to_3d(x, device).astype(np.float64)
    dx, ddx = derived_velocities(x, Ts)
    return dict(positions=positions, x=x, dx=dx, ddx=ddx, start_frame=start_frame, info=info)


def fte(DATA_DIR, start_frame, end_frame, dlc_thresh, fps=None, device=0, verbose=False):
    """Reference signature (all_optimizations.py:22): reads ``DATA_DIR/dlc/*.h5|csv`` and the scene
    file found above DATA_DIR, solves, and writes ``DATA_DIR/fte/fte.pickle``
    {positions, x, dx, ddx, start_frame}.  ``start_frame`` is 1-based like the reference's CLI;
    ``end_frame`` = -1 means the last frame.  fps is read from ``DATA_DIR/cam*.mp4`` when not given."""
    import os
    import pickle
    from glob import glob
    from time import time

    from . import utils

    t0 = time()
    assert os.path.exists(DATA_DIR)
    out_dir = os.path.join(DATA_DIR, "fte")
    dlc_dir = os.path.join(DATA_DIR, "dlc")
    assert os.path.exists(dlc_dir)
    os.makedirs(out_dir, exist_ok=True)
    paths = sorted(glob(os.path.join(dlc_dir, "*.h5"))) or sorted(glob(os.path.join(dlc_dir, "*.csv")))
    df = utils.load_dlc_points_as_df(paths, verbose=False)
    k_arr, d_arr, r_arr, t_arr, cam_res, n_cams, scene_fpath = utils.find_scene_file(DATA_DIR, verbose=False)
    tot_frames = int(df["frame"].max()) + 1
    if fps is None:
        import cv2

        vids = sorted(glob(os.path.join(DATA_DIR, "cam[1-9].mp4")))
        assert vids, "fps not given and no cam[1-9].mp4 next to the data"
        cap = cv2.VideoCapture(vids[0])
        fps = cap.get(cv2.CAP_PROP_FPS)
        cap.release()
    assert end_frame <= tot_frames, f"end_frame must be less than or equal to {tot_frames}"
    end_frame = tot_frames if end_frame == -1 else end_frame
    start_frame -= 1          # 0 based indexing (all_optimizations.py:59)
    assert start_frame >= 0
    print("\nInitialization took {0:.2f} seconds\n".format(time() - t0))
    t0 = time()
    out = fte_solve(df, k_arr, d_arr, r_arr, t_arr, start_frame, end_frame, dlc_thresh, fps, device=device, verbose=verbose)
    print("\nOptimization took {0:.2f} seconds\n".format(time() - t0))
    save_fte(out, out_dir, scene_fpath, start_frame, dlc_thresh, device=device)
    return out


# ---- output writers -----------------------------------------------------------
def save_3d_cheetah_as_2d(positions_3d_arr, out_dir, scene_fpath, bodyparts, project_func=None, start_frame=0,
                          save_as_csv=True, out_fname="fte", device=0):
    """Reprojection of an optimised trajectory into every camera of the scene, written as one
    DeepLabCut-style table per camera - the call at all_optimizations.py:560
    (``app.save_3d_cheetah_as_2d(positions, OUT_DIR, scene_fpath, markers, project_points_fisheye,
    start_frame)``; the callee lives in the reference's missing ``lib`` package, so the file format is the
    DLC layout its own loader reads back, utils.py:105-120: columns (scorer, bodyparts, x|y|likelihood),
    index = frame number).  All N x L points of a camera are projected by ONE kernel launch.  Markers whose 3-D
    position is not finite are written as NaN pixels with likelihood 0.
    Returns the list of per-camera DataFrames; files ``cam{i+1}_{out_fname}.h5`` (and ``.csv``)."""
    import os

    import pandas as pd

    from . import calib, utils

    k_arr, d_arr, r_arr, t_arr, _ = utils.load_scene(scene_fpath)
    P = np.asarray(positions_3d_arr, dtype=np.float64)
    N, L = P.shape[0], P.shape[1]
    assert L == len(bodyparts)
    project = calib.project_points_fisheye if project_func is None else project_func
    cols = pd.MultiIndex.from_product([["acinoset_b200"], list(bodyparts), ["x", "y", "likelihood"]],
                                      names=["scorer", "bodyparts", "coords"])
    index = np.arange(start_frame, start_frame + N)
    os.makedirs(out_dir, exist_ok=True)
    dfs = []
    for i in range(len(k_arr)):
        uv = np.asarray(project(P.reshape(-1, 3), k_arr[i], d_arr[i], r_arr[i], t_arr[i])).reshape(N, L, 2)
        # a marker without a 3-D estimate (NaN from TRI / SBA) has no reprojection: NaN pixels, likelihood 0 - never a
        # fabricated high-confidence detection at the projection of the world origin
        seen = np.isfinite(P).all(axis=2) & np.isfinite(uv).all(axis=2)
        uv = np.where(seen[..., None], uv, np.nan)
        data = np.concatenate([uv, seen[..., None].astype(np.float64)], axis=2).reshape(N, L * 3)
        df = pd.DataFrame(data, columns=cols, index=index)
        fpath = os.path.join(out_dir, f"cam{i + 1}_{out_fname}")
        try:
            df.to_hdf(fpath + ".h5", key="df_with_missing", format="table", mode="w")
        except ImportError:          # pytables missing: the csv twin carries the same table
            save_as_csv = True
        if save_as_csv:
            df.to_csv(fpath + ".csv")
        dfs.append(df)
    return dfs


def save_fte(out, out_dir, scene_fpath, start_frame, dlc_thresh, markers=None, device=0):
    """``app.save_fte`` at all_optimizations.py:559: the result pickle {positions, x, dx, ddx, start_frame}
    (:540-556) plus the 2-D reprojection files."""
    import os
    import pickle

    os.makedirs(out_dir, exist_ok=True)
    out_fpath = os.path.join(out_dir, "fte.pickle")
    with open(out_fpath, "wb") as f:
        pickle.dump({k: out[k] for k in ("positions", "x", "dx", "ddx", "start_frame")}, f)
    print(f"Saved {out_fpath}")
    save_3d_cheetah_as_2d(out["positions"], out_dir, scene_fpath, MARKERS if markers is None else markers, None,
                          start_frame, device=device)
    return out_fpath
