"""The reference's application layer by name (/root/reference/src/calib/app.py and the newer ``lib.app`` the pipeline
script calls, all_optimizations.py:20-975), for the functions that sit on the accelerated path.  GUI, plotting, video
and logging helpers (create_labeled_videos, plot_*, start_logging, ...) are out of scope.

    calibrate_fisheye_extrinsics_pairwise / calibrate_standard_extrinsics_pairwise    app.py:119-124
    sba_board_points_fisheye                                                          app.py:220-223
    sba_points_fisheye                                                                all_optimizations.py:874
    save_tri / save_sba / save_ekf / save_fte / save_optimised_cheetah / save_3d_cheetah_as_2d    :559-560,863,893,936
"""
import os
import pickle

from .fte import MARKERS, save_3d_cheetah_as_2d, save_fte  # noqa: F401
from .sba import sba_board_points, sba_board_points_fisheye, sba_points_fisheye  # noqa: F401
from .stereo import calibrate_fisheye_extrinsics_pairwise, calibrate_standard_extrinsics_pairwise  # noqa: F401


def save_optimised_cheetah(positions, out_fpath, extra_data=None):
    """all_optimizations.py:548-556 callee: pickle {positions, **extra_data}."""
    file_data = dict(positions=positions)
    if extra_data is not None:
        assert isinstance(extra_data, dict)
        file_data.update(extra_data)
    os.makedirs(os.path.dirname(os.path.abspath(out_fpath)), exist_ok=True)
    with open(out_fpath, "wb") as f:
        pickle.dump(file_data, f)
    print(f"Saved {out_fpath}")


def _save_stage(name, positions, out_dir, scene_fpath, start_frame, dlc_thresh, states=None, device=0):
    out_fpath = os.path.join(out_dir, f"{name}.pickle")
    extra = dict(start_frame=start_frame)
    if states is not None:
        extra.update(states)
    save_optimised_cheetah(positions, out_fpath, extra_data=extra)
    save_3d_cheetah_as_2d(positions, out_dir, scene_fpath, MARKERS, None, start_frame, out_fname=name, device=device)
    return out_fpath


def save_tri(positions, out_dir, scene_fpath, start_frame, dlc_thresh, device=0):
    """all_optimizations.py:936"""
    return _save_stage("tri", positions, out_dir, scene_fpath, start_frame, dlc_thresh, device=device)


def save_sba(positions, out_dir, scene_fpath, start_frame, dlc_thresh, device=0):
    """all_optimizations.py:893"""
    return _save_stage("sba", positions, out_dir, scene_fpath, start_frame, dlc_thresh, device=device)


def save_ekf(states, out_dir, scene_fpath, start_frame, dlc_thresh, device=0):
    """all_optimizations.py:863: states = dict(x, dx, ddx, smoothed_x, smoothed_dx, smoothed_ddx) in the EKF's state order."""
    from . import ekf

    positions = ekf.get_3d_marker_coords(states["smoothed_x"], device)
    return _save_stage("ekf", positions, out_dir, scene_fpath, start_frame, dlc_thresh, states=states, device=device)
