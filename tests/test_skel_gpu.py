"""GPU: the generic-skeleton FTE variant (reference src/build.py) through the C ABI vs the fp64 oracle."""
import json
import pickle

import numpy as np
import pytest

from conftest import golden
from oracle import skel_fte
from test_skel_host import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,loss", [("K1", "abs"), ("K1", "redescending"), ("K2", "abs")])
def test_skel_eval_matches_oracle(dummy_cams, tag, loss):
    from acinoset_b200 import fte

    skel, flat, x, meas, w = make_problem(tag, dummy_cams, 7, seed=3)
    K, D, R, t, _ = dummy_cams
    h = fte.set_scene(K, D, R, t)
    h.skel_set(flat, loss=loss)
    cost, g, Hu = h.skel_eval(x, meas, w)
    c0, g0, H0 = skel_fte.skel_eval(skel, x, meas, w, K, D, R, t, loss=loss)
    P = x.shape[1]
    assert np.abs(cost - c0).max() < 1e-9 * np.abs(c0).max()
    assert np.abs(g - g0).max() < 1e-9 * np.abs(g0).max()
    assert np.abs(skel_fte.upper_unpack(Hu, P) - H0).max() < 1e-9 * np.abs(H0).max()
    # cost / gradient only
    c1, g1, H1 = h.skel_eval(x, meas, w, want_H=False)
    assert H1 is None and np.array_equal(c1, cost) and np.array_equal(g1, g)
    # bit-reproducible
    c2, g2, Hu2 = h.skel_eval(x, meas, w)
    assert np.array_equal(Hu2, Hu) and np.array_equal(g2, g)


def test_skel_eval_requires_setup():
    from acinoset_b200 import _lib

    h = _lib.Handle(0)
    with pytest.raises(_lib.AcinoError):
        h.skel_set(dict(parts=["a"], link_parent=np.zeros(0, np.int32), link_flags=np.zeros(0, np.int32),
                        link_tv=np.zeros((0, 3)), out_order=[0], out_path=np.zeros(1, np.uint64),
                        dof_mask=np.zeros(1, np.int32)))          # cameras not set


def _smooth_truth(tag, N, rng):
    g = golden("generic_fk.npz")
    x = np.array(g[tag + "_x"][:N], dtype=np.float64)            # the reference's shipped (smooth) IPOPT solution
    t = np.arange(N) / 120.0
    x[:, 0] = 2.0 + 1.5 * t
    x[:, 1] = 6.5 + 0.8 * t
    x[:, 2] = 1.0 + 0.05 * np.sin(2 * np.pi * 2 * t)
    return x


def test_skel_solve_matches_oracle_lm(dummy_cams):
    """Same LM algorithm, fp64 on both sides: the GPU solve follows the CPU restatement."""
    from acinoset_b200 import build, fte, skeleton

    g = golden("generic_fk.npz")
    skel = json.loads(str(g["K1_skeleton_json"]))
    flat = skeleton.flatten_skeleton(skel)
    K, D, R, t, _ = dummy_cams
    rng = np.random.default_rng(21)
    N = 10
    x_true = _smooth_truth("K1", N, rng)
    f, names = skel_fte.pose_function(skel)
    P3 = np.array([f(r) for r in x_true])
    meas = np.stack([skel_fte.project(P3, K[c], D[c], R[c], t[c]) for c in range(len(K))], 1) + rng.normal(0, 1.0, (N, len(K), len(names), 2))
    w = np.where(rng.random(meas.shape[:-1]) < 0.9, 1.0 / skel_fte.MEAS_SIGMA_R, 0.0)
    x0 = x_true + rng.normal(0, 0.03, x_true.shape)
    handle = fte.set_scene(K, D, R, t)
    solver = build.SkelSolver(handle, flat, meas, w, 1 / 120.0, loss="abs")
    x_gpu, info = solver.solve(x0, max_iter=6)
    x_cpu, info_cpu = skel_fte.solve(skel, x0, meas, w, (K, D, R, t), 1 / 120.0, loss="abs", max_iter=6)
    assert info["F"] < info["F0"]
    assert abs(info["F"] - info_cpu["F"]) < 1e-6 * abs(info_cpu["F"])
    used = np.abs(x_cpu - x0).max(axis=0) > 0
    assert np.abs(x_gpu - x_cpu)[:, used].max() < 1e-4
    assert np.all(x_gpu[:, ~used] == x0[:, ~used])               # unused slots never move
    lo, hi = build.bounds(len(flat["parts"]))
    assert np.all(x_gpu[:-1] >= lo - 1e-12) and np.all(x_gpu[:-1] <= hi + 1e-12)


def test_build_model_and_solve_optimisation(tmp_path, dummy_cams):
    """build_model(skel_dict, project_dir) + solve_optimisation(...) on a synthetic project directory."""
    import pandas as pd
    from acinoset_b200 import build, utils

    g = golden("generic_fk.npz")
    skel = json.loads(str(g["K1_skeleton_json"]))
    K, D, R, t, res = dummy_cams
    K, D, R, t = K[:4], D[:4], R[:4], t[:4]
    rng = np.random.default_rng(8)
    N, start = 30, 5
    x_true = _smooth_truth("K1", N + start, rng)
    f, names = skel_fte.pose_function(skel)
    P3 = np.array([f(r) for r in x_true])
    uv = np.stack([skel_fte.project(P3, K[c], D[c], R[c], t[c]) for c in range(4)], 1) + rng.normal(0, 0.5, (N + start, 4, len(names), 2))
    (tmp_path / "data").mkdir()
    utils.save_scene(str(tmp_path / "data" / "4_cam_scene_static_sba.json"), K, D.reshape(-1, 4, 1), R, t.reshape(-1, 3, 1), res)
    cols = pd.MultiIndex.from_product([["scorer"], names, ["x", "y", "likelihood"]], names=["scorer", "bodyparts", "coords"])
    for c in range(4):
        arr = np.concatenate([uv[:, c], np.full((N + start, len(names), 1), 0.9)], axis=-1).reshape(N + start, -1)
        pd.DataFrame(arr, columns=cols).to_csv(tmp_path / "data" / f"cam{c + 1}DLC.csv")
    model, pose_to_3d = build.build_model(skel, str(tmp_path), N=N, start_frame=start, pair_by="name")
    assert model.meas.shape == (N, 4, len(names), 2) and model.x0.shape == (N, 3 + 3 * 15)
    assert np.all(model.w[:, :, names.index("neck")] == 0)        # "neck" is skipped (build.py:123-124)
    out = build.solve_optimisation(model, None, str(tmp_path), pose_to_3d, max_iter=60)
    with open(tmp_path / "data" / "results" / "traj_results.pickle", "rb") as fh:
        saved = pickle.load(fh)
    assert set(saved) == {"positions", "x", "dx", "ddx"}
    assert saved["positions"].shape == (N, len(names), 3) and saved["x"].shape == (N, 48)
    assert model.info["F"] < 0.2 * model.info["F0"]
    seen = [i for i, nm in enumerate(names) if nm != "neck"]
    err = np.linalg.norm(saved["positions"][:, seen] - P3[start:, seen], axis=-1)
    assert np.median(err) < 0.02, np.median(err)
    assert np.abs(saved["x"][1:] - saved["x"][:-1] - model.h * saved["dx"][1:]).max() < 1e-12
