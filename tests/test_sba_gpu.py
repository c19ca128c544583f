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


def _pinhole_cams():
    """Dummy 6-camera scene with OpenCV standard-model coefficients (rational + thin prism) instead of fisheye ones."""
    import synth

    K, _, R, t, res = synth.load_dummy_scene()
    rng = np.random.default_rng(11)
    D = np.zeros((len(K), 12))
    D[:, 0] = -0.10 + 0.01 * rng.normal(size=len(K))      # k1
    D[:, 1] = 0.02 + 0.005 * rng.normal(size=len(K))      # k2
    D[:, 2:4] = 1e-3 * rng.normal(size=(len(K), 2))       # p1 p2
    D[:, 4] = -0.002                                      # k3
    D[:, 5] = 0.01                                        # k4
    D[:, 8:12] = 1e-3 * rng.normal(size=(len(K), 4))      # s1..s4
    return K, D, R, t, res


def test_pinhole_model_residuals_and_jacobian_match_oracle():
    """app.sba_board_points (app.py:215-218) passes cv2.projectPoints / undistortPoints wrappers: with
    project_func=calib.project_points the SBA kernels evaluate the standard model - residuals against the oracle's
    restatement of cv2.projectPoints (pinned to cv2 in test_oracle_golden), the Jacobian against central differences."""
    import synth
    from acinoset_b200 import calib, sba
    from oracle import pinhole

    cams = _pinhole_cams()
    K, D = cams[0], cams[1]
    p = synth.make_sba_problem(12, pinhole.project_points, seed=4, cams=cams)
    n_pts = len(p["points_3d_true"])
    C = len(K)
    rv = np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]])
    x = np.concatenate([rv, p["t0"].ravel(), p["points_3d_true"].ravel() + 0.01])
    args = (C, n_pts, p["point_3d_indices"], p["camera_indices"], K, D, p["points_2d"])
    f = sba.cost_func_points_extrinsics(x, *args, project_func=calib.project_points)
    obj, r_arr, t_arr = sba.params_to_points_extrinsics(x, C, n_pts)
    ref = np.empty((len(p["point_3d_indices"]), 2))
    for c in range(C):
        m = p["camera_indices"] == c
        ref[m] = pinhole.project_points(obj[p["point_3d_indices"][m]], K[c], D[c], r_arr[c], t_arr[c]) - p["points_2d"][m].astype(np.float64)
    assert np.abs(f - ref.ravel()).max() < 1e-8
    # the fisheye default reads the same coefficients differently: the model switch is real
    f_fish = sba.cost_func_points_extrinsics(x, C, n_pts, p["point_3d_indices"], p["camera_indices"], K, D[:, :4], p["points_2d"])
    assert np.abs(f_fish - f).max() > 1.0
    J = sba.jac_points_extrinsics(x, *args, project_func=calib.project_points).toarray()
    rng = np.random.default_rng(0)
    for j in list(rng.integers(0, 6 * C, 8)) + list(rng.integers(6 * C, len(x), 8)):
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[j] += h
        xm[j] -= h
        fd = (sba.cost_func_points_extrinsics(xp, *args, project_func=calib.project_points) -
              sba.cost_func_points_extrinsics(xm, *args, project_func=calib.project_points)) / (2 * h)
        assert np.abs(J[:, j] - fd).max() < 1e-4 * max(1.0, np.abs(fd).max())
    with pytest.raises(NotImplementedError):
        sba.cost_func_points_extrinsics(x, *args, project_func=lambda *a: None)


def test_sba_board_points_pinhole_recovers_extrinsics(tmp_path):
    """sba_board_points by the reference's name (app.py:215-218): points files + scene file in, refined scene out; the
    perturbed extrinsics come back to the truth (relative poses, gauge free) and the cost drops to the noise floor."""
    import synth
    from acinoset_b200 import calib, sba, utils
    from oracle import pinhole

    cams = _pinhole_cams()
    K, D, R, t, res = cams
    C = len(K)
    p = synth.make_sba_problem(60, pinhole.project_points, seed=5, cams=cams, noise_px=0.1)
    n_pts = len(p["points_3d_true"])
    ppi = 54
    # points files in the reference's layout: per camera, the views it saw
    fpaths = []
    for c in range(C):
        m = p["camera_indices"] == c
        views = np.unique(p["point_3d_indices"][m] // ppi)
        pts = np.stack([p["points_2d"][m & (p["point_3d_indices"] // ppi == v)] for v in views]).reshape(len(views), ppi, 1, 2)
        fp = tmp_path / f"points_{c}.json"
        utils.save_points(str(fp), pts, [f"img{v:05d}.jpg" for v in views], (9, 6), 0.1, res)
        fpaths.append(str(fp))
    scene_in, scene_out = tmp_path / "scene.json", tmp_path / "scene_sba.json"
    utils.save_scene(str(scene_in), K, D, p["R0"], p["t0"].reshape(C, 3, 1), res)
    out = sba.sba_board_points(str(scene_in), fpaths, str(scene_out))
    assert set(out) == {"before", "after"}
    c0 = 0.5 * np.sum(np.log1p(out["before"] ** 2))
    c1 = 0.5 * np.sum(np.log1p(out["after"] ** 2))
    assert c1 < 0.05 * c0 and np.sqrt(np.mean(out["after"] ** 2)) < 0.2
    K2, D2, R2, t2, _ = utils.load_scene(str(scene_out))
    assert np.allclose(np.asarray(D2).reshape(C, -1)[:, :12], D)

    def rel(Ra, ta, Rb, tb):
        return Rb @ Ra.T, tb.reshape(3) - Rb @ Ra.T @ ta.reshape(3)

    # the global scale is part of the gauge (points and baselines scale together): fix it on the first baseline
    worst_r = worst_t = 0.0
    scale = None
    for c in range(1, C):
        Rr, tr = rel(R2[0], t2[0], R2[c], t2[c])
        Rt, tt = rel(p["R_true"][0], p["t_true"][0], p["R_true"][c], p["t_true"][c])
        scale = np.linalg.norm(tt) / np.linalg.norm(tr) if scale is None else scale
        worst_r = max(worst_r, np.degrees(np.arccos(np.clip((np.trace(Rr @ Rt.T) - 1) / 2, -1, 1))))
        worst_t = max(worst_t, np.linalg.norm(scale * tr - tt))
    assert worst_r < 0.05 and worst_t < 0.01 and abs(scale - 1) < 0.05, (worst_r, worst_t, scale)
    # the reference keeps the SBA functions in calib.py too: same objects by that name
    assert calib.bundle_adjust_board_points_and_extrinsics is sba.bundle_adjust_board_points_and_extrinsics


@pytest.mark.parametrize("tag", ["static", "rotating"])
def test_bundle_adjust_matches_independent_optimiser(tag):
    """K6 / K7 against SciPy TRF with the reference's options and the ANALYTIC Jacobian (tests/golden/solves.npz,
    tests/golden/make_golden_solves.py): same cost to 1e-7 relative, same relative pose of the two cameras (the gauge is
    free) to 1e-5.  The reference's own run stops earlier (2.2845e+01 / 5.3361e+01) because its finite-difference Jacobian
    is grouped by a sparsity pattern that does not match its parameter layout (test_reference_sparsity_pattern_misses_...)."""
    from acinoset_b200 import calib, sba

    g, s = golden("sba.npz"), golden("solves.npz")
    K, D, R, t = g[f"{tag}_K"], g[f"{tag}_D"], g[f"{tag}_R"], g[f"{tag}_t"]
    obj, r_new, t_new, res = calib.bundle_adjust_points_and_extrinsics(
        g[f"{tag}_points_2d"], g[f"{tag}_points_3d"], g[f"{tag}_pidx"], g[f"{tag}_cidx"], K, D, R, t, None)
    cost = 0.5 * np.sum(np.log1p(res["after"] ** 2))
    cost_ref = float(s[f"sba_{tag}_cost"])
    assert abs(cost - cost_ref) < 1e-7 * cost_ref, (cost, cost_ref)
    _, r_ref, t_ref = sba.params_to_points_extrinsics(s[f"sba_{tag}_x"], 2, 864)

    def rel(Ra, ta, Rb, tb):
        return Rb @ Ra.T, tb.reshape(3) - Rb @ Ra.T @ ta.reshape(3)

    Rr, tr = rel(r_new[0], t_new[0], r_new[1], t_new[1])
    Rs, ts = rel(r_ref[0], t_ref[0], r_ref[1], t_ref[1])
    # scale is part of the gauge: compare the direction and the length ratio of the baseline separately
    assert np.abs(Rr - Rs).max() < 1e-5, np.abs(Rr - Rs).max()
    assert np.abs(tr / np.linalg.norm(tr) - ts / np.linalg.norm(ts)).max() < 1e-5
