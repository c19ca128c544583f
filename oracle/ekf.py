"""CPU restatement (test infrastructure, NOT the product) of the reference's EKF measurement model and
filter loop: /root/reference/src/all_optimizations.py:615-649 (h_function, numerical_jacobian with
eps = 1e-3 forward differences) and :773-846 (filter + RTS smoother).  ``misc.get_3d_marker_coords`` lives
in the reference's missing ``lib`` package; it is the cheetah FK of :138-190 evaluated on the EKF's
joint-grouped state order (:734-746), restated here with oracle.skeleton.  Parity unpinned by reference
tests (there are none for the EKF); pinned against the analytic fp64 Jacobian of oracle.fte."""
import numpy as np

from . import fisheye, fte, skeleton

EKF_TO_ACTIVE = np.array([0, 1, 2, 3, 6, 20, 4, 7, 21, 8, 5, 9, 22, 10, 23, 11, 24, 12, 13, 14, 15, 16, 17, 18, 19])


def to_active(x_ekf):
    out = np.empty_like(np.asarray(x_ekf, dtype=np.float64))
    out[..., EKF_TO_ACTIVE] = x_ekf
    return out


def h_function(x, k, d, r, t):
    """:615-621 for one camera: (20,2)."""
    return fisheye.project(skeleton.cheetah_fk_active(to_active(x)[None])[0], k, d, r, t)


def numerical_jacobian(func, x, *args):
    """:634-649, literally (forward differences, eps = 1e-3)."""
    n = len(x)
    eps = 1e-3
    fx = func(x, *args).flatten()
    xp = x.copy()
    jac = np.empty((len(fx), n))
    for i in range(n):
        xp[i] = xp[i] + eps
        jac[:, i] = (func(xp, *args).flatten() - fx) / eps
        xp[i] = x[i]
    return jac


def measurement_numerical(x, K, D, R, t):
    """h (C*40,), H (C*40, 25) exactly as assembled at :797-806."""
    C = len(K)
    hs, Hs = [], []
    for j in range(C):
        hs.append(h_function(x, K[j], D[j], R[j], t[j]).flatten())
        Hs.append(numerical_jacobian(h_function, np.asarray(x, dtype=np.float64), K[j], D[j], R[j], t[j]))
    return np.concatenate(hs), np.concatenate(Hs)


def measurement_analytic(x, K, D, R, t):
    """Same quantities with the closed-form fp64 Jacobian (oracle.fte.residuals_and_jac)."""
    xa = to_active(np.asarray(x, dtype=np.float64))[None]
    C = len(K)
    r, J = fte.residuals_and_jac(xa, np.zeros((1, C, 20, 2)), K, D, R, t)
    return r[0].reshape(-1), J[0].reshape(-1, 25)[:, EKF_TO_ACTIVE]


def ekf_loop(pixels_arr, likelihood_arr, states0, fps, dlc_thresh, max_pixel_err, K, D, R, t, analytic=False):
    """:713-846 restated literally (per-frame loops, per-pair gating loop, numerical Jacobian)."""
    n = 25
    n_states = 75
    vel_idx, acc_idx = 25, 50
    sT = 1.0 / fps
    n_cams = len(K)
    n_markers = 20
    sigma_bound = 3
    p_lin_pos = np.ones(3) * 3 ** 2
    p_ang_pos = np.ones(n - 3) * (np.pi / 4) ** 2
    p_lin_vel = np.ones(3) * 5 ** 2
    p_ang_vel = np.ones(n - 3) * 3 ** 2
    p_lin_acc = np.ones(3) * 3 ** 2
    p_ang_acc = np.ones(n - 3) * 3 ** 2
    p_ang_acc[10:] = 5 ** 2
    P = np.diag(np.concatenate([p_lin_pos, p_ang_pos, p_lin_vel, p_ang_vel, p_lin_acc, p_ang_acc]))
    qb_list = [5.0, 5.0, 5.0, 10.0, 10.0, 10.0, 5.0, 25.0, 5.0, 50.0, 5.0, 50.0, 25.0, 100.0, 30.0, 140.0, 40.0,
               350.0, 200.0, 350.0, 200.0, 450.0, 400.0, 450.0, 400.0]
    qb = (np.diag(qb_list) / 2) ** 2
    Q = np.block([[sT ** 4 / 4 * qb, sT ** 3 / 2 * qb, sT ** 2 / 2 * qb],
                  [sT ** 3 / 2 * qb, sT ** 2 * qb, sT * qb],
                  [sT ** 2 / 2 * qb, sT * qb, qb]])
    dlc_cov = 5 ** 2
    rng = np.arange(n_states - vel_idx)
    rng_acc = np.arange(n_states - acc_idx)
    F = np.eye(n_states)
    F[rng, rng + vel_idx] = sT
    F[rng_acc, rng_acc + acc_idx] = sT ** 2 / 2
    n_frames = pixels_arr.shape[0]
    states = np.asarray(states0, dtype=np.float64).copy()
    states_est_hist = np.zeros((n_frames, n_states))
    states_pred_hist = states_est_hist.copy()
    P_est_hist = np.zeros((n_frames, n_states, n_states))
    P_pred_hist = P_est_hist.copy()
    outliers_ignored = 0
    for i in range(n_frames):
        acc_prediction = states[acc_idx:]
        vel_prediction = states[vel_idx:acc_idx] + sT * acc_prediction
        pos_prediction = states[:vel_idx] + sT * vel_prediction + (0.5 * sT ** 2) * acc_prediction
        states = np.concatenate([pos_prediction, vel_prediction, acc_prediction]).astype(np.float32).flatten()
        states_pred_hist[i] = states
        P = F @ P @ F.T + Q
        P_pred_hist[i] = P
        z_k = pixels_arr[i]
        likelihood = likelihood_arr[i]
        H = np.zeros((n_cams * n_markers * 2, n_states))
        if analytic:
            h, H[:, :vel_idx] = measurement_analytic(states[:vel_idx], K, D, R, t)
        else:
            h, H[:, :vel_idx] = measurement_numerical(states[:vel_idx].astype(np.float64), K, D, R, t)
        bad_point_mask = np.repeat(likelihood < dlc_thresh, 2)
        dlc_cov_arr = dlc_cov * np.ones((n_cams * n_markers * 2))
        dlc_cov_arr[bad_point_mask] = max_pixel_err
        Rm = np.diag(dlc_cov_arr ** 2)
        residual = z_k - h
        S = (H @ P @ H.T) + Rm
        temp = sigma_bound * np.sqrt(np.diag(S))
        for j in range(0, len(residual), 2):
            if np.abs(residual[j]) > temp[j] or np.abs(residual[j + 1]) > temp[j + 1]:
                residual[j:j + 2] = 0
                outliers_ignored += 1
        Kg = P @ H.T @ np.linalg.inv(S)
        states = states + Kg @ residual
        states_est_hist[i] = states
        P = (np.eye(Kg.shape[0]) - Kg @ H) @ P
        P_est_hist[i] = P
    smooth = states_est_hist.copy()
    smooth_P = P_est_hist.copy()
    for i in range(n_frames - 2, 0, -1):
        A = P_est_hist[i] @ F.T @ np.linalg.inv(P_pred_hist[i + 1])
        smooth[i] = states_est_hist[i] + A @ (smooth[i + 1] - states_pred_hist[i + 1])
        smooth_P[i] = P_est_hist[i] + A @ (smooth_P[i + 1] - P_pred_hist[i + 1]) @ A.T
    return dict(x=states_est_hist[:, :vel_idx], smoothed_x=smooth[:, :vel_idx], outliers_ignored=outliers_ignored)
