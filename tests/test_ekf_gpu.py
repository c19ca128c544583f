"""GPU: dense measurement Jacobian kernel (fte_jac) and the EKF built on it against the fp64 oracle and
the literal restatement of the reference's numerical-Jacobian EKF loop (all_optimizations.py:615-649,773-846)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle(dummy_cams):
    import acinoset_b200 as ab

    K, D, R, t, _ = dummy_cams
    h = ab.Handle(0)
    h.set_cameras(K, D, R, t)
    yield h
    h.close()


def _problem(n, seed=5):
    import synth
    from oracle import fisheye, skeleton

    cams = synth.load_dummy_scene()
    return synth.make_fte_problem(n, skeleton.cheetah_fk_active, fisheye.project, seed=seed, cams=cams), cams


@pytest.mark.parametrize("n", [1, 7, 8, 203])
def test_fte_jac_matches_oracle(handle, dummy_cams, n):
    from oracle import fte as ofte

    p, cams = _problem(n)
    K, D, R, t, _ = cams
    x = p["x0"].astype(np.float32)
    uv, J = handle.fte_jac(x)
    assert uv.shape == (n, 6, 20, 2) and J.shape == (n, 6, 20, 2, 25)
    r, Jo = ofte.residuals_and_jac(x.astype(np.float64), np.zeros((n, 6, 20, 2)), K, D, R, t)
    w = p["w"] > 0                                   # theta < 60 deg, inside the image (the stated tolerance domain)
    assert np.abs(uv - r)[w].max() < 2e-3            # px
    num = np.linalg.norm((J - Jo)[w].reshape(-1, 25), axis=1)
    den = np.linalg.norm(Jo[w].reshape(-1, 25), axis=1)
    assert (num / den).max() < 1e-4
    # structural zeros are exact zeros
    assert np.all(J[Jo == 0] == 0)


def test_fte_jac_dev_equals_host_and_partial_outputs(handle, dummy_cams):
    import torch

    p, _ = _problem(19)
    x = p["x0"].astype(np.float32)
    uv, J = handle.fte_jac(x)
    xd = torch.from_numpy(x).cuda()
    Jd = torch.empty(19, 6, 20, 2, 25, device="cuda")
    handle.fte_jac_dev(xd, None, Jd)
    torch.cuda.synchronize()
    assert np.array_equal(Jd.cpu().numpy(), J)
    uv2, none = handle.fte_jac(x, want_J=False)
    assert none is None and np.array_equal(uv2, uv)
    _, uv_fk = handle.fk_project(x)
    assert np.abs(uv_fk - uv).max() < 1e-3


def test_measurement_jacobian_ekf_order_vs_reference_numerical(handle, dummy_cams):
    """Analytic H in the EKF's joint-grouped order vs the reference's forward-difference H (eps 1e-3)."""
    from acinoset_b200 import ekf, fte
    from oracle import ekf as oekf

    p, cams = _problem(3, seed=9)
    K, D, R, t, _ = cams
    fte.set_scene(K, D, R, t)
    x_ekf = ekf.from_active(p["x_true"][1])
    assert np.array_equal(ekf.to_active(x_ekf), p["x_true"][1])
    h, H = ekf.measurement_jacobian(x_ekf)
    ho, Ho = oekf.measurement_numerical(x_ekf, K, D, R, t)
    ha, Ha = oekf.measurement_analytic(x_ekf, K, D, R, t)
    w = np.repeat(p["w"][1].reshape(-1) > 0, 2)
    assert np.abs(h - ho)[w].max() < 2e-3
    assert np.abs(H - Ha)[w].max() < 1e-4 * np.abs(Ha[w]).max()
    # forward differences with eps = 1e-3 carry O(eps) truncation error: ~1e-3 relative
    assert np.abs(H - Ho)[w].max() < 5e-3 * np.abs(Ho[w]).max()
    # single-camera h_function keeps the reference signature
    h0 = ekf.h_function(x_ekf, K[0], D[0], R[0], t[0])
    assert h0.shape == (20, 2) and np.abs(h0.reshape(-1) - ho[:40])[w[:40]].max() < 2e-3


def test_ekf_filter_matches_reference_loop(handle, dummy_cams):
    import synth
    from acinoset_b200 import ekf, fte
    from oracle import ekf as oekf, fisheye, skeleton

    # the synthetic FTE trajectory slowed 4x and without gross outliers: the reference's EKF gates at 3 sigma of a
    # 25 px measurement std only after its covariance has contracted, so uniform-in-image outliers during the
    # first frames make it diverge (and a diverged filter is chaotic - nothing to compare)
    n = 40
    K, D, R, t, res = dummy_cams
    rng = np.random.default_rng(12)
    x_true = synth.make_trajectory(n, rng, fps=synth.FPS * 4)
    meas, lik = synth.make_measurements(skeleton.cheetah_fk_active(x_true), dummy_cams, fisheye.project, rng, outlier_frac=0.0)
    fte.set_scene(K, D, R, t)
    Ts = 1.0 / synth.FPS
    pixels = meas.reshape(n, -1).astype(np.float64)
    likf = lik.reshape(n, -1)
    states0 = np.zeros(75)
    states0[:25] = ekf.from_active(x_true[0])
    states0[25:50] = ekf.from_active((x_true[1] - x_true[0]) / Ts)
    out = ekf.ekf_filter(pixels, likf, states0, 1.0 / Ts, 0.5, res[0])
    ref_a = oekf.ekf_loop(pixels, likf, states0, 1.0 / Ts, 0.5, res[0], K, D, R, t, analytic=True)
    ref_n = oekf.ekf_loop(pixels, likf, states0, 1.0 / Ts, 0.5, res[0], K, D, R, t, analytic=False)
    # same algorithm, fp32 kernel Jacobian vs fp64 analytic: tight; vs the reference's eps = 1e-3 forward-difference
    # Jacobian: its truncation error
    assert out["outliers_ignored"] == ref_a["outliers_ignored"]
    assert np.abs(out["x"] - ref_a["x"]).max() < 2e-4
    assert np.abs(out["smoothed_x"] - ref_a["smoothed_x"]).max() < 2e-4
    assert np.abs(out["smoothed_x"] - ref_n["smoothed_x"]).max() < 5e-3
    # and the filter tracks the truth
    assert np.abs(ekf.to_active(out["smoothed_x"])[5:, :3] - x_true[5:, :3]).max() < 0.01
