"""GPU: the reference-level FTE entry points (DataFrame in, result dict / pickle out)."""
import json
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(N, cams):
    import synth
    from oracle import fisheye, skeleton

    return synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=31, cams=cams)


def test_fte_solve_from_dataframe(dummy_cams):
    import synth
    from acinoset_b200 import fte
    from oracle import skeleton

    K, D, R, t, _ = dummy_cams
    p = _problem(150, dummy_cams)
    df = synth.dense_to_long_df(p["meas"], p["lik"], fte.MARKERS)
    out = fte.fte_solve(df, K, D.reshape(-1, 4, 1), R, t.reshape(-1, 3, 1), 20, 140, 0.5, 120.0, max_iter=60)
    assert out["positions"].shape == (120, 20, 3) and out["x"].shape == (120, 25)
    assert out["dx"].shape == (120, 25) and out["ddx"].shape == (120, 25) and out["start_frame"] == 20
    Pt = skeleton.cheetah_fk_active(p["x_true"][20:140])
    rms = np.sqrt(((out["positions"] - Pt) ** 2).sum(-1).mean())
    assert rms < 0.02, rms                      # from the reference's crude linear initialisation
    # collocation identities of the saved states (K3)
    Ts = 1 / 120.0
    x, dx, ddx = out["x"], out["dx"], out["ddx"]
    assert np.abs(x[1:] - x[:-1] - Ts * dx[1:]).max() < 1e-12
    assert np.abs(dx[1:] - dx[:-1] - Ts * ddx[1:]).max() < 1e-9
    # positions are the FK of the saved states
    assert np.abs(out["positions"] - skeleton.cheetah_fk_active(x)).max() < 5e-6


def test_fte_reference_signature_writes_pickle(tmp_path, dummy_cams):
    import pandas as pd
    from acinoset_b200 import fte, utils

    K, D, R, t, res = dummy_cams
    p = _problem(60, dummy_cams)
    data = tmp_path / "run"
    (data / "dlc").mkdir(parents=True)
    (tmp_path / "extrinsic_calib").mkdir()
    utils.save_scene(str(tmp_path / "extrinsic_calib" / "6_cam_scene_sba.json"), K, D.reshape(-1, 4, 1), R, t.reshape(-1, 3, 1), res)
    cols = pd.MultiIndex.from_product([["scorer"], fte.MARKERS, ["x", "y", "likelihood"]], names=["scorer", "bodyparts", "coords"])
    for c in range(6):
        arr = np.concatenate([p["meas"][:, c], p["lik"][:, c, :, None]], axis=-1).reshape(60, -1)
        pd.DataFrame(arr, columns=cols).to_csv(data / "dlc" / f"cam{c + 1}DLC.csv")
    out = fte.fte(str(data), 1, -1, 0.5, fps=120.0)
    with open(data / "fte" / "fte.pickle", "rb") as f:
        saved = pickle.load(f)
    assert set(saved) == {"positions", "x", "dx", "ddx", "start_frame"}
    assert saved["positions"].shape == (60, 20, 3) and saved["x"].shape == (60, 25) and saved["start_frame"] == 0
    assert np.array_equal(saved["x"], out["x"])
    # 2-D reprojection files (all_optimizations.py:560): one DLC-style table per camera, read back by the loader
    from oracle import fisheye

    files = sorted((data / "fte").glob("cam*_fte.csv"))
    assert len(files) == 6
    df2 = utils.load_dlc_points_as_df([str(f) for f in files], verbose=False)
    assert len(df2) == 60 * 6 * 20
    dense, lik = utils.dlc_df_to_dense(df2, 6, fte.MARKERS, 0, 60)
    for c in range(6):
        uv = fisheye.project(saved["positions"], K[c], D[c], R[c], t[c])
        ok = np.abs(uv) < 1e4                       # behind-camera points project "validly" far outside the image
        assert np.abs(dense[:, c] - uv)[ok].max() < 1e-6 * np.abs(uv[ok]).max() + 1e-6
    assert np.all(lik == 1.0)
