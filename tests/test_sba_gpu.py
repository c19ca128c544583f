"""GPU: SBA residuals / analytic Jacobian / solve behind the reference's calib names, against the
golden vectors produced by the reference's own code and the fp64 oracle."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["static", "rotating"])
def test_cost_func_matches_reference_residuals(tag):
    from acinoset_b200 import calib

    g = golden("sba.npz")
    n_pts = len(g[f"{tag}_points_3d"])
    f = calib.cost_func_points_extrinsics(g[f"{tag}_x0"], 2, n_pts, g[f"{tag}_pidx"], g[f"{tag}_cidx"], g[f"{tag}_K"],
                                          g[f"{tag}_D"], g[f"{tag}_points_2d"], None)
    assert f.shape == g[f"{tag}_f0"].shape
    assert np.abs(f - g[f"{tag}_f0"]).max() < 1e-8
    cost = 0.5 * np.sum(np.log1p(f ** 2))
    assert f"{cost:.4e}" == {"static": "2.3636e+01", "rotating": "5.4156e+01"}[tag]      # notebook trace (K5)
    # points-only residuals at the same point
    R, t = g[f"{tag}_R"], g[f"{tag}_t"]
    f2 = calib.cost_func_points_only(g[f"{tag}_x0"][12:], n_pts, g[f"{tag}_pidx"], g[f"{tag}_cidx"], g[f"{tag}_K"],
                                     g[f"{tag}_D"], R, t, g[f"{tag}_points_2d"], None)
    assert np.abs(f2 - g[f"{tag}_f0"]).max() < 1e-8


def test_analytic_jacobian_matches_oracle_and_fd():
    from acinoset_b200 import sba
    from oracle import sba as osba

    g = golden("sba.npz")
    tag = "static"
    K, D = g[f"{tag}_K"], g[f"{tag}_D"].reshape(-1, 4)
    sel = np.concatenate([np.arange(30), 54 + np.arange(30)])
    pidx, cidx, p2d = g[f"{tag}_pidx"][sel], g[f"{tag}_cidx"][sel], g[f"{tag}_points_2d"][sel]
    n_pts = int(pidx.max()) + 1
    x0 = np.concatenate([g[f"{tag}_x0"][:12], g[f"{tag}_x0"][12:12 + 3 * n_pts]])
    J = sba.jac_points_extrinsics(x0, 2, n_pts, pidx, cidx, K, D, p2d).toarray()
    Jr, Jt, Jx = osba.jac_blocks_points_extrinsics(x0, 2, n_pts, pidx, cidx, K, D)
    Jo = osba.dense_jacobian(Jr, Jt, Jx, 2, n_pts, pidx, cidx)
    assert np.abs(J - Jo).max() < 1e-9 * np.abs(Jo).max()
    eps = 1e-6
    for k in (0, 4, 7, 11, 12, 20):
        d = np.zeros_like(x0)
        d[k] = eps
        fd = (sba.cost_func_points_extrinsics(x0 + d, 2, n_pts, pidx, cidx, K, D, p2d)
              - sba.cost_func_points_extrinsics(x0 - d, 2, n_pts, pidx, cidx, K, D, p2d)) / (2 * eps)
        assert np.abs(J[:, k] - fd).max() < 1e-5 * max(1.0, np.abs(fd).max())


def test_prepare_board_data_matches_oracle():
    from acinoset_b200 import calib
    from oracle import sba as osba

    g = golden("sba.npz")
    tag = "static"
    K, D, R, t = g[f"{tag}_K"], g[f"{tag}_D"], g[f"{tag}_R"], g[f"{tag}_t"]
    pts = [g[f"{tag}_img_pts_a"], g[f"{tag}_img_pts_b"]]
    fns = [list(g[f"{tag}_fnames_a"]), list(g[f"{tag}_fnames_b"])]
    p2, p3, pi, ci = calib.prepare_calib_board_data_for_bundle_adjustment(pts, fns, tuple(g["board_shape"]), K, D, R, t, None)
    o2, o3, oi, oc, _ = osba.prepare_calib_board_data(pts, fns, tuple(g["board_shape"]), K, D, R, t)
    assert p2.dtype == np.float32 and p3.dtype == np.float32
    assert np.array_equal(p2, o2) and np.array_equal(pi, oi) and np.array_equal(ci, oc)
    assert np.abs(p3 - o3).max() < 1e-5
    assert p2.shape == (1728, 2) and p3.shape == (864, 3)


@pytest.mark.parametrize("tag,ref_final", [("static", 2.2845e+01), ("rotating", 5.3361e+01)])
def test_bundle_adjust_extrinsics_reaches_reference_cost(tag, ref_final):
    """K6/K7: the shipped solutions stop (xtol) at cost 2.2845e+01 / 5.3361e+01 with first-order
    optimality ~1e3-1e4, i.e. not stationary; the GPU solve must end at or below that cost."""
    from acinoset_b200 import calib

    g = golden("sba.npz")
    K, D, R, t = g[f"{tag}_K"], g[f"{tag}_D"], g[f"{tag}_R"], g[f"{tag}_t"]
    obj, r_new, t_new, res = calib.bundle_adjust_points_and_extrinsics(
        g[f"{tag}_points_2d"], g[f"{tag}_points_3d"], g[f"{tag}_pidx"], g[f"{tag}_cidx"], K, D, R, t, None)
    assert obj.shape == (864, 3) and r_new.shape == (2, 3, 3) and t_new.shape == (2, 3, 1)
    assert set(res) == {"before", "after"} and res["before"].shape == (3456,) and res["after"].shape == (3456,)
    assert np.abs(res["before"] - g[f"{tag}_f0"]).max() < 1e-8
    cost_after = 0.5 * np.sum(np.log1p(res["after"] ** 2))
    assert cost_after <= ref_final * (1 + 1e-4)
    for c in range(2):
        assert np.abs(r_new[c] @ r_new[c].T - np.eye(3)).max() < 1e-12
    # stays in the neighbourhood of the shipped solution (gauge is free: loose bound)
    assert np.abs(r_new - g[f"{tag}_R_sba"]).max() < 5e-3 and np.abs(t_new - g[f"{tag}_t_sba"]).max() < 5e-2


def test_points_only_and_synthetic_six_camera_scene():
    import synth
    from acinoset_b200 import calib, sba
    from oracle import fisheye

    p = synth.make_sba_problem(120, fisheye.project, seed=2)
    K, D = p["K"], p["D"]
    n_pts = len(p["points_3d_true"])
    # cost at the truth (noise floor)
    prob = sba.SBAProblem(p["points_2d"], p["point_3d_indices"], p["camera_indices"], K, D, n_pts)
    x_true = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R_true"]]), p["t_true"].ravel()])
    f_true = prob.residuals(x_true, p["points_3d_true"])
    cost_true = 0.5 * np.sum(np.log1p(f_true ** 2))
    # initial points: two-view triangulation with the PERTURBED extrinsics (like the reference's assembly)
    pts0 = np.zeros((n_pts, 3))
    first = {}
    for i, (pi, ci) in enumerate(zip(p["point_3d_indices"], p["camera_indices"])):
        first.setdefault(pi, []).append(i)
    a_idx = np.array([v[0] for v in first.values()])
    b_idx = np.array([v[1] for v in first.values()])
    for (ca, cb) in set(zip(p["camera_indices"][a_idx], p["camera_indices"][b_idx])):
        m = (p["camera_indices"][a_idx] == ca) & (p["camera_indices"][b_idx] == cb)
        X = calib.triangulate_points_fisheye(p["points_2d"][a_idx[m]], p["points_2d"][b_idx[m]], K[ca], D[ca], p["R0"][ca],
                                             p["t0"][ca], K[cb], D[cb], p["R0"][cb], p["t0"][cb])
        pts0[p["point_3d_indices"][a_idx[m]]] = X
    obj, r_new, t_new, res, info = calib.bundle_adjust_points_and_extrinsics(
        p["points_2d"], pts0.astype(np.float32), p["point_3d_indices"], p["camera_indices"], K, D, p["R0"],
        p["t0"].reshape(-1, 3, 1), None, return_info=True)
    assert info["info"] == 0
    assert info["cost"] < info["cost0"] * 0.2 and info["cost"] < cost_true * 1.02
    # points-only refinement with the true cameras
    obj2, res2 = calib.bundle_adjust_points_only(p["points_2d"], pts0.astype(np.float32), p["point_3d_indices"],
                                                 p["camera_indices"], K, D, p["R_true"], p["t_true"].reshape(-1, 3, 1), None)
    assert np.sqrt(((obj2 - p["points_3d_true"]) ** 2).sum(-1).mean()) < 5e-3
    assert np.sum(res2["after"] ** 2) < np.sum(res2["before"] ** 2)
