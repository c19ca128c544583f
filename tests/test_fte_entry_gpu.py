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


def _write_run(tmp_path, p, cams, n):
    import pandas as pd
    from acinoset_b200 import fte, utils

    K, D, R, t, res = cams
    data = tmp_path / "run"
    (data / "dlc").mkdir(parents=True)
    (tmp_path / "extrinsic_calib").mkdir()
    utils.save_scene(str(tmp_path / "extrinsic_calib" / "6_cam_scene_sba.json"), K, D.reshape(-1, 4, 1), R, t.reshape(-1, 3, 1), res)
    cols = pd.MultiIndex.from_product([["scorer"], fte.MARKERS, ["x", "y", "likelihood"]], names=["scorer", "bodyparts", "coords"])
    for c in range(6):
        arr = np.concatenate([p["meas"][:, c], p["lik"][:, c, :, None]], axis=-1).reshape(n, -1)
        pd.DataFrame(arr, columns=cols).to_csv(data / "dlc" / f"cam{c + 1}DLC.csv")
    return data


def test_pipeline_tri_sba_ekf_by_reference_names(tmp_path, dummy_cams):
    """all_optimizations.tri / sba / ekf (:569-936) on a synthetic run directory."""
    import synth
    from acinoset_b200 import all_optimizations as ao
    from oracle import fisheye, skeleton

    n = 40
    rng = np.random.default_rng(5)
    x_true = synth.make_trajectory(n, rng, fps=synth.FPS * 4)
    P = skeleton.cheetah_fk_active(x_true)
    meas, lik = synth.make_measurements(P, dummy_cams, fisheye.project, rng, noise_px=0.5, outlier_frac=0.0)
    data = _write_run(tmp_path, dict(meas=meas, lik=lik), dummy_cams, n)
    pos = ao.tri(str(data), 1, -1, 0.5)
    assert pos.shape == (n, 20, 3)
    seen = ~np.isnan(pos[..., 0])
    assert seen.mean() > 0.5
    err_tri = np.linalg.norm(pos[seen] - P[seen], axis=-1)
    assert np.median(err_tri) < 0.02
    with open(data / "tri" / "tri.pickle", "rb") as f:
        saved = pickle.load(f)
    assert np.array_equal(np.isnan(saved["positions"]), np.isnan(pos)) and saved["start_frame"] == 0
    # 2-D reprojection files: a marker TRI could not triangulate is NaN with likelihood 0, never a confident detection
    # at the projection of the world origin; every other marker has likelihood 1
    import pandas as pd

    assert (~seen).any()
    for c in range(6):
        df = pd.read_csv(data / "tri" / f"cam{c + 1}_tri.csv", header=[0, 1, 2], index_col=0)
        lk = df.xs("likelihood", axis=1, level=2).to_numpy()
        xs = df.xs("x", axis=1, level=2).to_numpy()
        assert np.array_equal(lk == 1.0, np.isfinite(xs)) and np.all((lk == 0.0) | (lk == 1.0))
        assert np.all(lk[~seen] == 0.0) and np.all(np.isnan(xs[~seen]))
    pos_sba, residuals = ao.sba(str(data), 1, -1, 0.5)
    assert pos_sba.shape == (n, 20, 3) and set(residuals) == {"before", "after"}
    err_sba = np.linalg.norm(pos_sba[seen] - P[seen], axis=-1)
    # refining each point against every camera that saw it cannot be worse than the adjacent-pair mean
    assert 0.5 * np.sum(np.log1p((residuals["after"] / 50) ** 2)) <= 0.5 * np.sum(np.log1p((residuals["before"] / 50) ** 2)) + 1e-9
    assert np.median(err_sba) <= np.median(err_tri) * 1.05
    out = ao.ekf(str(data), 1, -1, 0.5, fps=synth.FPS)
    assert out["positions"].shape == (n, 20, 3) and out["smoothed_x"].shape == (n, 25) and out["start_frame"] == 0
    assert os.path.exists(data / "ekf" / "ekf.pickle") and len(list((data / "ekf").glob("cam*_ekf.csv"))) == 6


def test_pt3d_to_2d_scalar_and_array(dummy_cams):
    from acinoset_b200 import all_optimizations as ao
    from oracle import fisheye

    K, D, R, t, _ = dummy_cams
    X = np.array([[1.0, 6.0, 0.5], [2.5, 7.0, 0.2], [0.3, 5.0, 1.0]])
    ref = fisheye.project(X, K[1], D[1], R[1], t[1])
    u, v = ao.pt3d_to_2d(X[0, 0], X[0, 1], X[0, 2], K[1], D[1], R[1], t[1])
    assert isinstance(u, float) and abs(u - ref[0, 0]) < 1e-9 and abs(v - ref[0, 1]) < 1e-9
    uu, vv = ao.pt3d_to_2d(X[:, 0], X[:, 1], X[:, 2], K[1], D[1], R[1], t[1])
    assert np.abs(uu - ref[:, 0]).max() < 1e-9 and np.abs(vv - ref[:, 1]).max() < 1e-9
    assert abs(ao.pt3d_to_x2d(*X[2], K[1], D[1], R[1], t[1]) - ref[2, 0]) < 1e-9
    assert abs(ao.pt3d_to_y2d(*X[2], K[1], D[1], R[1], t[1]) - ref[2, 1]) < 1e-9
