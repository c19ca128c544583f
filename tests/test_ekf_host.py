"""CPU: the EKF oracle restatement (numerical vs analytic Jacobian, filter loop) and the host-side pieces of
acinoset_b200.ekf that need no GPU (state-order maps, F / Q / P0 of all_optimizations.py:713-764)."""
import numpy as np


def _setup(n, slow=4, seed=12, outlier_frac=0.0):
    import synth
    from oracle import fisheye, skeleton

    cams = synth.load_dummy_scene()
    rng = np.random.default_rng(seed)
    x_true = synth.make_trajectory(n, rng, fps=synth.FPS * slow)
    meas, lik = synth.make_measurements(skeleton.cheetah_fk_active(x_true), cams, fisheye.project, rng, outlier_frac=outlier_frac)
    return cams, x_true, meas, lik


def test_state_order_maps_and_matrices():
    from acinoset_b200 import ekf
    from oracle import ekf as oekf

    assert np.array_equal(ekf.EKF_TO_ACTIVE, oekf.EKF_TO_ACTIVE)
    assert sorted(ekf.EKF_TO_ACTIVE.tolist()) == list(range(25))
    x = np.arange(25.0)
    assert np.array_equal(ekf.from_active(ekf.to_active(x)), x)
    idx = ekf.get_pose_params()
    # the slots the reference initialises by name (all_optimizations.py:710-711)
    assert (idx["x_0"], idx["y_0"], idx["psi_0"]) == (0, 1, 5) and len(idx) == 25
    # psi_0 of the EKF order is the yaw slot (20) of the library's active order
    assert ekf.EKF_TO_ACTIVE[idx["psi_0"]] == 20
    sT = 1 / 120.0
    F = ekf.transition_matrix(sT)
    s = np.random.default_rng(0).normal(size=75)
    # the reference's prediction uses the PREDICTED velocity in the position update (:627-629), so it differs from
    # F s by dt^2 * acc in the position block - a quirk kept on both sides; velocity / acceleration blocks agree
    pred = ekf.predict_next_state(s, sT)
    Fs = F @ s
    assert pred.dtype == np.float32                                             # the cast at :631
    assert np.allclose(pred[25:], Fs[25:], rtol=1e-6, atol=1e-6)
    assert np.allclose(pred[:25], Fs[:25] + sT ** 2 * s[50:], rtol=1e-6, atol=1e-6)
    Q = ekf.process_covariance(sT)
    assert Q.shape == (75, 75) and np.allclose(Q, Q.T) and np.linalg.eigvalsh(Q).min() > -1e-9
    assert np.isclose(Q[74, 74], (400.0 / 2) ** 2) and np.isclose(Q[0, 0], sT ** 4 / 4 * 2.5 ** 2)
    P0 = ekf.initial_covariance()
    assert np.isclose(P0[0, 0], 9.0) and np.isclose(P0[3, 3], (np.pi / 4) ** 2) and np.isclose(P0[74, 74], 25.0) and np.isclose(P0[50 + 3 + 9, 50 + 3 + 9], 9.0)


def test_oracle_numerical_vs_analytic_jacobian():
    from oracle import ekf as oekf

    cams, x_true, meas, lik = _setup(3)
    K, D, R, t, _ = cams
    x = x_true[1][oekf.EKF_TO_ACTIVE]
    h, H = oekf.measurement_numerical(x, K, D, R, t)
    ha, Ha = oekf.measurement_analytic(x, K, D, R, t)
    ok = np.repeat(lik[1].reshape(-1) > 0.5, 2)
    assert np.abs(h - ha).max() < 1e-9
    # forward differences with eps = 1e-3 (all_optimizations.py:639): O(eps) truncation error
    assert np.abs(H - Ha)[ok].max() < 2e-3 * np.abs(Ha[ok]).max()
    # structure: each marker depends on 6..17 of the 25 parameters
    nz = (np.abs(Ha.reshape(6, 20, 2, 25)[0, :, 0, :]) > 0).sum(axis=1)
    assert nz.tolist() == [6, 6, 6, 9, 10, 13, 15, 17, 10, 11, 12, 10, 11, 12, 13, 14, 15, 13, 14, 15]


def test_oracle_ekf_loop_tracks_truth():
    from oracle import ekf as oekf

    n = 16
    cams, x_true, meas, lik = _setup(n)
    K, D, R, t, res = cams
    Ts = 1 / 120.0
    xe = x_true[:, oekf.EKF_TO_ACTIVE]
    s0 = np.zeros(75)
    s0[:25] = xe[0]
    s0[25:50] = (xe[1] - xe[0]) / Ts
    out = oekf.ekf_loop(meas.reshape(n, -1).astype(np.float64), lik.reshape(n, -1), s0, 1 / Ts, 0.5, res[0], K, D, R, t, analytic=True)
    assert out["outliers_ignored"] == 0
    assert np.abs(out["smoothed_x"][3:, :3] - xe[3:, :3]).max() < 0.01
