"""CPU-only: the C-ABI library loads and exports every symbol include/acino_b200.h declares;
host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    return True


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "acino_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acino_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from acinoset_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/acino_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == declared
    assert lib.acino_version() >= 100


def test_no_gpu_fails_loudly(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import acinoset_b200 as ab

    with pytest.raises(ab.AcinoError):
        ab.Handle(0)


def test_new_entry_points_have_no_cpu_fallback(built):
    """The generic-skeleton solve and the pair calibration need the GPU like everything else."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import json

    import acinoset_b200 as ab
    from acinoset_b200 import build, stereo

    g = np.load(os.path.join(ROOT, "tests", "golden", "stereo.npz"))
    with pytest.raises(ab.AcinoError):
        stereo.calibrate_pair_extrinsics_fisheye(g["rot12_obj"], g["rot12_img1"], g["rot12_img2"], g["rot12_K1"], g["rot12_D1"],
                                                 g["rot12_K2"], g["rot12_D2"], (1920, 1080))
    gk = np.load(os.path.join(ROOT, "tests", "golden", "generic_fk.npz"))
    skel = json.loads(str(gk["K1_skeleton_json"]))
    K = np.tile(np.eye(3), (2, 1, 1))
    model = build.model_from_arrays(skel, (K, np.zeros((2, 4)), K.copy(), np.zeros((2, 3))), np.zeros((4, 2, 15, 2)), np.ones((4, 2, 15)))
    with pytest.raises(ab.AcinoError):
        build.solve_optimisation(model)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "acinoset_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt.replace("/root/reference/src", "REF").replace("(/root/reference", "(REF"), f


def test_scene_loaders_roundtrip(tmp_path):
    from acinoset_b200 import utils

    k, d, r, t, res = utils.load_scene(os.path.join(ROOT, "acinoset_b200", "data", "dummy_scene.json"))
    assert k.shape == (6, 3, 3) and d.shape == (6, 4, 1) and r.shape == (6, 3, 3) and t.shape == (6, 3, 1)
    assert res == (2704, 1520)
    out = tmp_path / "scene.json"
    utils.save_scene(str(out), k, d, r, t, res)
    k2, d2, r2, t2, res2 = utils.load_scene(str(out))
    assert np.array_equal(k, k2) and np.array_equal(d, d2) and np.array_equal(r, r2) and np.array_equal(t, t2)
    k3, d3, r3, t3, res3, n, path = utils.find_scene_file(str(tmp_path), scene_fname="scene.json", verbose=False)
    assert n == 6 and path == str(out)


def test_points_loader_reads_both_key_dialects(tmp_path):
    import json

    from acinoset_b200 import utils

    pts = np.arange(2 * 54 * 2, dtype=np.float32).reshape(2, 54, 2)
    for key, ts in (("board_edge_len", "created_timestamp"), ("board_square_len", "timestamp")):
        f = tmp_path / f"{key}.json"
        json.dump({ts: "x", "board_shape": [9, 6], key: 0.031, "camera_resolution": [2704, 1520],
                   "points": {"0.png": pts[0].tolist(), "1.png": pts[1].tolist()}}, open(f, "w"))
        p, fn, shape, edge, res = utils.load_points(str(f))
        assert np.array_equal(p, pts) and fn == ["0.png", "1.png"] and shape == (9, 6) and edge == 0.031
    utils.save_points(str(tmp_path / "o.json"), pts, ["a", "b"], (9, 6), 0.1, (10, 20))
    p, fn, shape, edge, res = utils.load_points(str(tmp_path / "o.json"))
    assert np.array_equal(p, pts) and edge == 0.1 and res == (10, 20)
    obj = utils.create_board_object_pts((9, 6), 0.5)
    assert obj.shape == (54, 3) and obj[1, 0] == 0.5 and obj[9, 1] == 0.5 and obj[:, 2].max() == 0


def test_dlc_tables_to_dense(tmp_path):
    import pandas as pd

    from acinoset_b200 import utils

    cols = pd.MultiIndex.from_product([["scorer"], ["nose", "spine"], ["x", "y", "likelihood"]],
                                      names=["scorer", "bodyparts", "coords"])
    paths = []
    for c in range(2):
        tab = pd.DataFrame(np.arange(5 * 6, dtype=float).reshape(5, 6) + 100 * c, columns=cols)
        tab.iloc[:, 2] = 0.9
        tab.iloc[:, 5] = 0.1 * (c + 1)
        f = tmp_path / f"cam{c}.csv"
        tab.to_csv(f)
        paths.append(str(f))
    df = utils.load_dlc_points_as_df(paths, verbose=False)
    assert list(df.columns) == ["frame", "camera", "marker", "x", "y", "likelihood"] and len(df) == 20
    meas, lik = utils.dlc_df_to_dense(df, 2, ["nose", "spine"], start_frame=1, n_frames=3)
    assert meas.shape == (3, 2, 2, 2) and lik.shape == (3, 2, 2)
    # frame 1, camera 1, spine = row 1 of the second table: x = 6*1+3+100, y = +4
    assert meas[0, 1, 1, 0] == 109 and meas[0, 1, 1, 1] == 110 and abs(lik[0, 1, 1] - 0.2) < 1e-6
    assert abs(lik[2, 0, 0] - 0.9) < 1e-6


def test_rotation_conversions_match_cv2_golden():
    from conftest import golden
    from acinoset_b200 import rotations

    g = golden("fisheye.npz")
    for rv, Rm, rb in zip(g["rvec"], g["rmat"], g["rvec_back"]):
        assert np.abs(rotations.rodrigues_to_mat(rv) - Rm).max() < 1e-14
        assert np.abs(rotations.rodrigues_to_vec(Rm) - rb).max() < 1e-12


def test_generic_skeleton_flattening_matches_oracle_walk():
    """Host logic of the generic builder (no GPU): the flattened link table reproduces the oracle's
    step-by-step replay of build.py:32-95 on random states."""
    import json

    from conftest import golden
    from acinoset_b200 import skeleton
    from oracle import skeleton as osk

    g = golden("generic_fk.npz")
    for tag in ("K1", "K2"):
        skel = json.loads(str(g[tag + "_skeleton_json"]))
        flat = skeleton.flatten_skeleton(skel)
        assert flat["out_names"] == list(g[tag + "_pose_order"])
        P = len(flat["parts"])
        rng = np.random.default_rng(3)
        x = rng.normal(0, 0.7, 3 + 3 * P)
        ref, names = osk.generic_fk_builder(skel)(x)
        # replay the flat table on the host with the oracle's rotation helpers
        pose = {i: x[:3].copy() for i in range(P)}
        for a, b, fl, tv in zip(flat["link_parent"], flat["link_child"], flat["link_flags"], flat["link_tv"]):
            L = np.eye(3)
            m = flat["dof_mask"][a]
            if m & 2:
                L = osk.rot_y(x[3 + P + a]) @ L
            if m & 1:
                L = osk.rot_x(x[3 + a]) @ L
            if m & 4:
                L = osk.rot_z(x[3 + 2 * P + a]) @ L
            pose[b] = pose[a] + (L.T if fl else L) @ tv
        out = np.stack([pose[i] for i in flat["out_order"]])
        assert np.abs(out - ref).max() < 1e-13
