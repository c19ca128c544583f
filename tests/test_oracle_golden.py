"""Pin the oracle (oracle/*.py) against golden vectors produced by the reference's own
code (tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from conftest import golden
from oracle import fisheye, fte, loss, sba, skeleton, triangulate


# ---------------------------------------------------------------- cheetah FK (K-FK)
def test_cheetah_fk_matches_reference_lambdify():
    g = golden("cheetah_fk.npz")
    assert list(g["active"]) == list(skeleton.ACTIVE_IDX)
    pos = skeleton.cheetah_fk(g["x45"])
    assert np.abs(pos - g["positions"]).max() < 1e-13


def test_cheetah_fk_jacobian_matches_sympy():
    g = golden("cheetah_fk.npz")
    xa = g["x45"][:12][:, skeleton.ACTIVE_IDX]
    J = skeleton.cheetah_fk_jac(xa)
    assert np.abs(J - g["jac"]).max() < 1e-12
    # sparsity of d positions / d state: 556 non-zeros of 60 x 25
    assert int((np.abs(g["jac"]).max(axis=0) > 0).sum()) == 556


def test_cheetah_fk_jacobian_matches_fd():
    rng = np.random.default_rng(5)
    xa = rng.normal(0, 0.7, (6, 25))
    J = skeleton.cheetah_fk_jac(xa)
    Jfd = skeleton.cheetah_fk_jac_fd(xa)
    assert np.abs(J - Jfd).max() < 1e-8


# ---------------------------------------------------------------- generic builder (K1, K2)
@pytest.mark.parametrize("tag", ["K1", "K2"])
def test_generic_fk_builder_known_answers(tag):
    g = golden("generic_fk.npz")
    skel = json.loads(str(g[tag + "_skeleton_json"]))
    f = skeleton.generic_fk_builder(skel)
    x = g[tag + "_x"]
    pos, order = f(x[0])
    assert order == list(g[tag + "_pose_order"])
    allpos = np.array([f(row)[0] for row in x])
    assert np.abs(allpos - g[tag + "_positions_pickle"]).max() < 1e-12
    assert np.abs(allpos - g[tag + "_positions_exec"]).max() < 1e-12


def test_generic_fk_dof_override_matters():
    # K2 note of SURVEY section 4: using new_human.pickle (markers non-empty => dofs forced to
    # [1,1,1]) on run1's states gives a 2.7e-4 mismatch -> the override at build.py:37-38 matters
    g = golden("generic_fk.npz")
    skel = json.loads(str(g["K1_skeleton_json"]))
    f = skeleton.generic_fk_builder(skel)
    x = g["K2_x"]
    allpos = np.array([f(row)[0] for row in x])
    d = np.abs(allpos - g["K2_positions_pickle"]).max()
    assert 1e-5 < d < 1e-2


# ---------------------------------------------------------------- collocation (K3)
@pytest.mark.parametrize("tag", ["K1", "K2"])
def test_collocation_identities(tag):
    g = golden("generic_fk.npz")
    x, dx, ddx = g[tag + "_x"], g[tag + "_dx"], g[tag + "_ddx"]
    h = 1 / 120  # build.py:131
    assert np.abs(x[1:] - x[:-1] - h * dx[1:]).max() < 1e-13
    assert np.abs(dx[1:] - dx[:-1] - h * ddx[1:]).max() < 1e-13
    d3 = (x[3:] - 3 * x[2:-1] + 3 * x[1:-2] - x[:-3]) / h ** 2
    slack = ddx[3:] - ddx[2:-1]
    assert np.abs(slack - d3).max() < 1e-9
    # and our derived dx / ddx reproduce the pickle away from the free leading values
    dxo, ddxo = fte.derived_velocities(x, h)
    assert np.abs(dxo[1:] - dx[1:]).max() < 1e-9
    assert np.abs(ddxo[2:] - ddx[2:]).max() < 1e-6


# ---------------------------------------------------------------- fisheye (K4) + cv2 pins
def test_projection_matches_pt3d_to_2d_and_cv2():
    g = golden("fisheye.npz")
    for c in range(6):
        uv = fisheye.project(g["X"], g["K"][c], g["D"][c], g["R"][c], g["t"][c])
        assert np.abs(uv - g["uv_pt3d_to_2d"][c]).max() < 1e-10
        assert np.abs(uv - g["uv_cv2"][c]).max() < 1e-9  # K4: 1.2e-10 px model difference


def test_projection_jacobian_fd():
    g = golden("fisheye.npz")
    X = g["X"]
    for c in (0, 3):
        cam = (g["K"][c], g["D"][c], g["R"][c], g["t"][c])
        uv, Jw, Jc = fisheye.project_jac(X, *cam)
        eps = 1e-6
        for k in range(3):
            d = np.zeros(3)
            d[k] = eps
            fd = (fisheye.project(X + d, *cam) - fisheye.project(X - d, *cam)) / (2 * eps)
            assert np.abs(Jw[..., k] - fd).max() < 1e-6 * max(1.0, np.abs(fd).max())


def test_undistort_matches_cv2_including_sentinel():
    g = golden("fisheye.npz")
    und = fisheye.undistort(g["und_in"], g["K"][1], g["D"][1])
    assert np.abs(und - g["und_dummy"]).max() < 1e-12 * max(1, np.abs(g["und_dummy"]).max())
    und = fisheye.undistort(g["und_in"], g["K_kiara"], g["D_kiara"])
    ref = g["und_kiara"]
    sent = ref[:, 0] == -1e6
    assert sent.sum() > 0 and (~sent).sum() > 0
    assert np.array_equal(und[sent], ref[sent])
    assert np.abs(und[~sent] - ref[~sent]).max() < 1e-9 * max(1, np.abs(ref[~sent]).max())


def test_rodrigues_matches_cv2():
    g = golden("fisheye.npz")
    for rv, Rm, Rj, rb in zip(g["rvec"], g["rmat"], g["rjac"], g["rvec_back"]):
        assert np.abs(fisheye.rodrigues(rv) - Rm).max() < 1e-14
        assert np.abs(fisheye.rodrigues_inv(Rm) - rb).max() < 1e-12
        if np.linalg.norm(rv) > 1e-6:
            d = fisheye.drodrigues(rv)  # [i,j,k]
            assert np.abs(d - Rj.reshape(3, 3, 3).transpose(1, 2, 0)).max() < 1e-12


# ---------------------------------------------------------------- loss
def test_redescending_loss_matches_reference():
    g = golden("loss.npz")
    assert np.abs(loss.redescending_loss(g["e"]) - g["rho_3_10_20"]).max() < 1e-12
    assert np.abs(loss.redescending_loss(g["e"], 3, 5, 15) - g["rho_3_5_15"]).max() < 1e-12
    assert abs(loss.redescending_loss(1e3) - 40.5) < 1e-9  # a*b - a^2/2 + a(c-b)/2


def test_redescending_derivatives_fd():
    e = np.linspace(0.01, 45, 1200)
    rho, d, dd = loss.redescending_dloss(e)
    h = 1e-6
    fd = (loss.redescending_loss(e + h) - loss.redescending_loss(e - h)) / (2 * h)
    assert np.abs(d - fd).max() < 1e-7
    fd2 = (loss.redescending_dloss(e + h)[1] - loss.redescending_dloss(e - h)[1]) / (2 * h)
    assert np.abs(dd - fd2).max() < 1e-6
    # odd symmetry of the derivative, weight is even and non-negative
    assert np.allclose(loss.redescending_dloss(-e)[1], -d)
    w = loss.redescending_irls_weight(np.concatenate([-e, [0.0], e]))
    assert (w >= 0).all() and abs(w[len(e)] - (1 - loss.func_step(3.0, 0.0))) < 1e-15


def test_cauchy_reproduces_notebook_initial_costs():
    # K5: calib_with_gui.ipynb:1415,1474 print initial cost 5.4156e+01 / 2.3636e+01
    g = golden("sba.npz")
    assert f"{loss.cauchy_cost(g['static_f0']):.4e}" == "2.3636e+01"
    assert f"{loss.cauchy_cost(g['rotating_f0']):.4e}" == "5.4156e+01"


# ---------------------------------------------------------------- triangulation / TRI (config 1, K8)
def test_pair_triangulation_matches_reference():
    g = golden("triangulate.npz")
    f = golden("fisheye.npz")
    K, D, R, t = f["K"], f["D"], f["R"], f["t"]
    out = triangulate.triangulate_points_fisheye(g["pair_a"], g["pair_b"], K[2], D[2], R[2], t[2], K[3], D[3], R[3], t[3])
    ref = g["pair_out"]
    scale = np.maximum(1.0, np.abs(ref))
    assert (np.abs(out - ref) / scale).max() < 1e-8


def test_tri_driver_config1_matches_reference():
    g = golden("triangulate.npz")
    f = golden("fisheye.npz")
    K, D, R, t = f["K"], f["D"].reshape(-1, 4), f["R"], f["t"].reshape(-1, 3)
    pos, cnt = triangulate.pairwise_mean_dense(g["meas"], g["lik"] > 0.5, K, D, R, t)
    ref = g["tri_pos"]
    assert np.array_equal(np.isnan(pos[..., 0]), np.isnan(ref[..., 0]))
    ok = ~np.isnan(ref[..., 0])
    scale = np.maximum(1.0, np.abs(ref[ok]))
    assert (np.abs(pos[ok] - ref[ok]) / scale).max() < 1e-8
    # K8: noise-free input is recovered exactly where a pair sees the point
    pos0, _ = triangulate.pairwise_mean_dense(g["meas_clean"], g["lik_clean"] > 0.5, K, D, R, t)
    ok0 = ~np.isnan(g["tri_pos_clean"][..., 0])
    assert np.abs(pos0[ok0] - g["P_true"][ok0]).max() < 1e-10


# ---------------------------------------------------------------- SBA assembly / residuals / Jacobian
@pytest.mark.parametrize("tag", ["static", "rotating"])
def test_sba_assembly_and_residuals_match_reference(tag):
    g = golden("sba.npz")
    K, D, R, t = g[f"{tag}_K"], g[f"{tag}_D"], g[f"{tag}_R"], g[f"{tag}_t"]
    pts = [g[f"{tag}_img_pts_a"], g[f"{tag}_img_pts_b"]]
    fns = [list(g[f"{tag}_fnames_a"]), list(g[f"{tag}_fnames_b"])]
    # visit views in the order the reference happened to use (it iterates a set)
    p2d, p3d, pidx, cidx, views = sba.prepare_calib_board_data(pts, fns, tuple(g["board_shape"]), K, D, R, t)
    # the reference's view order is arbitrary; compare as sets of per-view blocks via costs
    n_cam, n_pts = 2, len(p3d)
    x0 = sba.pack_params(R, t, p3d)
    f0 = sba.cost_func_points_extrinsics(x0, n_cam, n_pts, pidx, cidx, K, D.reshape(-1, 4), p2d)
    assert f0.shape == g[f"{tag}_f0"].shape
    # (points_3d is cast to float32 at calib.py:260; a last-bit difference between our DLT and
    # cv2's flips a few of those roundings => 1e-4 px residual changes => ~1e-5 in the cost)
    assert abs(loss.cauchy_cost(f0) - float(g[f"{tag}_cost0"])) < 2e-4
    # with the reference's own ordering the residual vector matches element-wise
    f0b = sba.cost_func_points_extrinsics(g[f"{tag}_x0"], n_cam, n_pts, g[f"{tag}_pidx"], g[f"{tag}_cidx"], K,
                                          D.reshape(-1, 4), g[f"{tag}_points_2d"])
    assert np.abs(f0b - g[f"{tag}_f0"]).max() < 1e-8
    # sparsity pattern
    A = sba.sparsity(n_cam, 6, g[f"{tag}_cidx"], n_pts, g[f"{tag}_pidx"])
    rows, cols = np.nonzero(A)
    assert tuple(g[f"{tag}_A_shape"]) == A.shape
    assert np.array_equal(rows, g[f"{tag}_A_rows"]) and np.array_equal(cols, g[f"{tag}_A_cols"])


def test_sba_analytic_jacobian_fd():
    g = golden("sba.npz")
    tag = "static"
    K, D = g[f"{tag}_K"], g[f"{tag}_D"].reshape(-1, 4)
    pidx, cidx, p2d = g[f"{tag}_pidx"][:40], g[f"{tag}_cidx"][:40], g[f"{tag}_points_2d"][:40]
    # small sub-problem: first 40 observations of camera 0 + matching ones of camera 1
    sel = np.concatenate([np.arange(40), 54 + np.arange(40)])
    pidx, cidx, p2d = g[f"{tag}_pidx"][sel], g[f"{tag}_cidx"][sel], g[f"{tag}_points_2d"][sel]
    n_pts = int(pidx.max()) + 1
    x0 = np.concatenate([g[f"{tag}_x0"][:12], g[f"{tag}_x0"][12:12 + 3 * n_pts]])
    Jr, Jt, Jx = sba.jac_blocks_points_extrinsics(x0, 2, n_pts, pidx, cidx, K, D)
    J = sba.dense_jacobian(Jr, Jt, Jx, 2, n_pts, pidx, cidx)
    eps = 1e-6
    Jfd = np.zeros_like(J)
    for k in range(len(x0)):
        d = np.zeros_like(x0)
        d[k] = eps
        Jfd[:, k] = (sba.cost_func_points_extrinsics(x0 + d, 2, n_pts, pidx, cidx, K, D, p2d)
                     - sba.cost_func_points_extrinsics(x0 - d, 2, n_pts, pidx, cidx, K, D, p2d)) / (2 * eps)
    assert np.abs(J - Jfd).max() < 1e-5 * np.abs(Jfd).max()


# ---------------------------------------------------------------- FTE objective
def test_fte_eval_gradient_fd(fte_problem_small):
    p = fte_problem_small
    K, D, R, t, _ = p["cams"]
    xa = p["x0"][:6]
    cost, g, H = fte.fte_eval(xa, p["meas"][:6], p["w"][:6], K, D, R, t)
    eps = 1e-6
    for k in range(25):
        d = np.zeros(25)
        d[k] = eps
        cp, _, _ = fte.fte_eval(xa + d, p["meas"][:6], p["w"][:6], K, D, R, t)
        cm, _, _ = fte.fte_eval(xa - d, p["meas"][:6], p["w"][:6], K, D, R, t)
        fd = (cp - cm) / (2 * eps)
        assert np.abs(fd - g[:, k]).max() < 1e-5 * max(1.0, np.abs(g[:, k]).max())
    # H symmetric PSD
    assert np.abs(H - np.swapaxes(H, 1, 2)).max() < 1e-9 * np.abs(H).max()
    assert np.linalg.eigvalsh(H).min() > -1e-6 * np.abs(H).max()
    assert np.allclose(fte.unpack_upper(fte.pack_upper(H)), H)


def test_smooth_term_gradient_and_band():
    rng = np.random.default_rng(2)
    xa = rng.normal(0, 0.1, (12, 25))
    Ts = 1 / 120
    g = fte.smooth_grad(xa, Ts)
    eps = 1e-3  # the term is quadratic: central differences are exact up to rounding
    for (n, p) in [(0, 0), (3, 7), (5, 20), (11, 24), (8, 12)]:
        d = np.zeros_like(xa)
        d[n, p] = eps
        fd = (fte.smooth_cost(xa + d, Ts) - fte.smooth_cost(xa - d, Ts)) / (2 * eps)
        assert abs(fd - g[n, p]) < 1e-7 * max(1.0, abs(g[n, p]))
    band = fte.smooth_band(12, Ts)
    # band . x reproduces the gradient (quadratic form)
    gb = np.zeros_like(xa)
    for n in range(12):
        for k in range(4):
            if n + k < 12:
                gb[n] += band[n, k] * xa[n + k]
                if k:
                    gb[n + k] += band[n, k] * xa[n]
    assert np.abs(gb - g).max() < 1e-9 * np.abs(g).max()
    assert fte.smooth_cost(xa[:3], Ts) == 0.0


@pytest.mark.parametrize("tag", ["d5", "d8", "d12"])
def test_pinhole_twins_match_reference(tag):
    """project_points / create_undistort_point_function / triangulate_points (calib.py:25-30,52-66)."""
    from oracle import pinhole

    g = golden("pinhole.npz")
    d = g[tag]
    uv1 = pinhole.project_points(g["X"], g["K1"], d, g["r1"], g["t1"])
    assert np.abs(uv1 - g[f"{tag}_uv1"]).max() < 1e-9
    assert np.abs(uv1 - g[f"{tag}_uv1_rvec"]).max() < 1e-9
    uv2 = pinhole.project_points(g["X"], g["K2"], d, g["r2"], g["t2"])
    assert np.abs(uv2 - g[f"{tag}_uv2"]).max() < 1e-9
    und = pinhole.undistort_points(g[f"{tag}_uv1"], g["K1"], d, to_pixels=True)
    assert np.abs(und - g[f"{tag}_und"]).max() < 1e-9
    tri = pinhole.triangulate_points(g[f"{tag}_noisy1"], g[f"{tag}_noisy2"], g["K1"], d, g["r1"], g["t1"],
                                     g["K2"], d, g["r2"], g["t2"])
    assert np.abs(tri - g[f"{tag}_tri"]).max() < 1e-8


@pytest.mark.parametrize("N,seed", [(48, 5), (100, 6)])
def test_lm_restatement_reaches_independent_optimum(N, seed):
    """The LM algorithm the CUDA path runs (oracle/lm.py, fp64) against an independent optimiser (SciPy trust-exact on the
    same objective from the same start, tests/golden/solves.npz): objective at least as low to 1e-6, markers within 1e-3 m."""
    import synth
    from oracle import fisheye, fte as ofte, lm as olm, skeleton

    g = golden("solves.npz")
    cams = synth.load_dummy_scene()
    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=seed, cams=cams)
    x, info = olm.solve(p, p["x0"], max_iter=60, eval_dtype=np.float32)
    K, D, R, t, _ = cams
    meas, w = p["meas"].astype(np.float32).astype(np.float64), p["w"].astype(np.float32).astype(np.float64)
    c, _, _ = ofte.fte_eval(x, meas, w, K, D, R, t)
    F = float(c.sum()) + ofte.smooth_cost(x, p["Ts"], ofte.model_weights_active())
    assert F <= float(g[f"fte{N}_F"]) * (1 + 1e-6)
    assert np.abs(skeleton.cheetah_fk_active(x) - skeleton.cheetah_fk_active(g[f"fte{N}_x"])).max() < 1e-3


def test_reference_sparsity_pattern_misses_true_nonzeros():
    """The reference hands SciPy a camera-major sparsity pattern (calib.py:196-207) for a parameter vector that is laid out
    [all rvecs | all tvecs | points] (calib.py:346-351): a third of the analytic Jacobian's non-zeros fall OUTSIDE the
    pattern, so its grouped finite differences are wrong and its runs stop at a non-stationary point (the shipped
    4_cam_scene_*_sba.json).  Quantified on the K6 / K7 fixtures at the reference's own starting point."""
    from oracle import sba as osba

    g, s = golden("sba.npz"), golden("solves.npz")
    for tag in ("static", "rotating"):
        K, D = g[f"{tag}_K"], g[f"{tag}_D"].reshape(-1, 4)
        pidx, cidx = g[f"{tag}_pidx"], g[f"{tag}_cidx"]
        n_pts = len(g[f"{tag}_points_3d"])
        Jr, Jt, Jx = osba.jac_blocks_points_extrinsics(g[f"{tag}_x0"], 2, n_pts, pidx, cidx, K, D)
        J = osba.dense_jacobian(Jr, Jt, Jx, 2, n_pts, pidx, cidx)
        A = osba.sparsity(2, 6, cidx, n_pts, pidx)
        nz = np.abs(J) > 1e-12
        outside = int(np.count_nonzero(nz & (A == 0)))
        assert outside == int(s[f"sba_{tag}_outside"]) == 10368 and int(nz.sum()) == 31104
        # with the true layout every non-zero is inside: the pattern is right for a camera-major vector only
        cam_major = np.concatenate([np.r_[3 * c:3 * c + 3, 6 + 3 * c:6 + 3 * c + 3] for c in range(2)])
        J_cm = np.concatenate([J[:, cam_major], J[:, 12:]], axis=1)
        assert np.count_nonzero((np.abs(J_cm) > 1e-12) & (A == 0)) == 0
        # and the independent optimum is below what the reference's run reached
        assert float(s[f"sba_{tag}_cost"]) < {"static": 2.2845e+01, "rotating": 5.3361e+01}[tag]
