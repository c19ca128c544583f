"""Extended Kalman filter + RTS smoother over the cheetah pose - host side.

Mirrors the EKF section of the reference (src/all_optimizations.py:569-866): constant-acceleration
prediction (:624-631, :759-764), measurement h(x) = project_points_fisheye(FK(x)) for every camera
(:615-621, :800-806), 3-sigma residual gating (:819-824), Kalman update (:826-834) and the
Rauch-Tung-Striebel smoother (:839-846).  The filter recursion is sequential in time and stays on the
CPU (NumPy, like the reference); what moves to the GPU is its hot spot: the reference evaluates h
26 times per camera per frame (numerical_jacobian, :634-649, eps = 1e-3, one cv2 call each) - here ONE
launch of fte_jac_kernel (csrc/fte_jac.cu) returns h and the analytic Jacobian for all cameras.

State order (the EKF's joint-grouped 25 pose parameters, qb_list at :734-746):
    x,y,z, phi0,theta0,psi0, phi1,theta1,psi1, theta2, phi3,theta3,psi3, theta4,psi4, theta5,psi5,
    theta6..theta13
followed by the 25 velocities and 25 accelerations.
"""
import numpy as np

from . import _lib
from . import fte as _fte

N_POSE = _lib.N_ACTIVE
N_MARKERS = _lib.N_MARKERS

POSE_PARAMS = ["x_0", "y_0", "z_0", "phi_0", "theta_0", "psi_0", "phi_1", "theta_1", "psi_1", "theta_2",
               "phi_3", "theta_3", "psi_3", "theta_4", "psi_4", "theta_5", "psi_5", "theta_6", "theta_7",
               "theta_8", "theta_9", "theta_10", "theta_11", "theta_12", "theta_13"]
# EKF slot -> slot of the library's active order [x,y,z, phi0,phi1,phi3, theta0..13, psi0,psi1,psi3,psi4,psi5]
EKF_TO_ACTIVE = np.array([0, 1, 2, 3, 6, 20, 4, 7, 21, 8, 5, 9, 22, 10, 23, 11, 24, 12, 13, 14, 15, 16, 17, 18, 19])
# process noise std-devs qb_list (all_optimizations.py:734-746)
QB_LIST = np.array([5.0, 5.0, 5.0, 10.0, 10.0, 10.0, 5.0, 25.0, 5.0, 50.0, 5.0, 50.0, 25.0, 100.0, 30.0, 140.0, 40.0,
                    350.0, 200.0, 350.0, 200.0, 450.0, 400.0, 450.0, 400.0])


def get_pose_params():
    """misc.get_pose_params() as used at all_optimizations.py:582-592: name -> state index."""
    return {name: i for i, name in enumerate(POSE_PARAMS)}


def to_active(x_ekf):
    """EKF-ordered pose (…,25) -> the library's active order."""
    x_ekf = np.asarray(x_ekf)
    out = np.empty_like(x_ekf)
    out[..., EKF_TO_ACTIVE] = x_ekf
    return out


def from_active(x_active):
    return np.asarray(x_active)[..., EKF_TO_ACTIVE]


def get_3d_marker_coords(x, device=0):
    """misc.get_3d_marker_coords (all_optimizations.py:618): EKF-ordered pose -> (20,3) marker positions."""
    return _fte.pose_to_3d(to_active(np.asarray(x, dtype=np.float64)[..., :N_POSE]), device).astype(np.float64)


def h_function(x, k, d, r, t, device=0):
    """all_optimizations.py:615-621: (20,2) pixels of the markers in one camera."""
    from . import calib

    return calib.project_points_fisheye(get_3d_marker_coords(x, device), k, d, r, t)


def measurement_jacobian(x, device=0):
    """h (C*20*2,) and H (C*20*2, 25) for ALL cameras of the installed scene (set_scene) in one kernel
    launch - the analytic replacement of the loop at all_optimizations.py:800-806.  x: EKF-ordered
    pose (25,) or a batch (N,25) -> (N, C*40), (N, C*40, 25)."""
    x = np.asarray(x, dtype=np.float64)
    single = x.ndim == 1
    xa = to_active(np.atleast_2d(x)[:, :N_POSE]).astype(np.float32)
    uv, J = _fte.get_handle(device).fte_jac(xa)
    N = xa.shape[0]
    h = uv.reshape(N, -1).astype(np.float64)
    H = J.reshape(N, -1, N_POSE)[:, :, EKF_TO_ACTIVE].astype(np.float64)
    return (h[0], H[0]) if single else (h, H)


def predict_next_state(x, dt):
    """all_optimizations.py:624-631 (including its float32 cast)."""
    n = x.shape[0] // 3
    acc = x[2 * n:]
    vel = x[n:2 * n] + dt * acc
    pos = x[:n] + dt * vel + (0.5 * dt ** 2) * acc
    return np.concatenate([pos, vel, acc]).astype(np.float32)


def initial_covariance():
    """P0 of all_optimizations.py:713-732."""
    n = N_POSE
    p_ang_acc = np.ones(n - 3) * 3 ** 2
    p_ang_acc[10:] = 5 ** 2
    return np.diag(np.concatenate([np.ones(3) * 3 ** 2, np.ones(n - 3) * (np.pi / 4) ** 2, np.ones(3) * 5 ** 2,
                                   np.ones(n - 3) * 3 ** 2, np.ones(3) * 3 ** 2, p_ang_acc]))


def process_covariance(sT):
    """Q of all_optimizations.py:748-753."""
    qb = (np.diag(QB_LIST) / 2) ** 2
    return np.block([[sT ** 4 / 4 * qb, sT ** 3 / 2 * qb, sT ** 2 / 2 * qb],
                     [sT ** 3 / 2 * qb, sT ** 2 * qb, sT * qb],
                     [sT ** 2 / 2 * qb, sT * qb, qb]])


def transition_matrix(sT):
    """F of all_optimizations.py:759-764."""
    n = N_POSE
    F = np.eye(3 * n)
    r = np.arange(2 * n)
    F[r, r + n] = sT
    ra = np.arange(n)
    F[ra, ra + 2 * n] = sT ** 2 / 2
    return F


def ekf_filter(pixels_arr, likelihood_arr, states0, fps, dlc_thresh, max_pixel_err, device=0, jacobian=None,
               sigma_bound=3, dlc_cov=5 ** 2):
    """The filter + smoother loops of all_optimizations.py:773-846 for frames 0..n-1.

    pixels_arr (n, C*20*2) and likelihood_arr (n, C*20) in the reference's [camera][marker][x,y] column
    order; states0 (75,) initial state; the scene must have been installed with fte.set_scene.
    jacobian(x_pose) -> (h, H) defaults to the GPU kernel.  Returns the dict the reference saves
    (x, dx, ddx, smoothed_x, smoothed_dx, smoothed_ddx) plus outliers_ignored."""
    jac = (lambda xp: measurement_jacobian(xp, device)) if jacobian is None else jacobian
    n_frames = pixels_arr.shape[0]
    n = N_POSE
    n_states = 3 * n
    sT = 1.0 / fps
    states = np.asarray(states0, dtype=np.float64).copy()
    P = initial_covariance()
    Q = process_covariance(sT)
    F = transition_matrix(sT)
    est = np.zeros((n_frames, n_states))
    pred = est.copy()
    P_est = np.zeros((n_frames, n_states, n_states))
    P_pred = P_est.copy()
    outliers_ignored = 0
    n_meas = pixels_arr.shape[1]
    for i in range(n_frames):
        states = predict_next_state(states, sT).flatten()
        pred[i] = states
        P = F @ P @ F.T + Q
        P_pred[i] = P
        z_k = pixels_arr[i]
        h, Hp = jac(states[:n])
        H = np.zeros((n_meas, n_states))
        H[:, :n] = Hp
        bad = np.repeat(likelihood_arr[i] < dlc_thresh, 2)
        cov = dlc_cov * np.ones(n_meas)
        cov[bad] = max_pixel_err
        R = np.diag(cov ** 2)
        residual = z_k - h
        S = (H @ P @ H.T) + R
        temp = sigma_bound * np.sqrt(np.diag(S))
        out = (np.abs(residual[0::2]) > temp[0::2]) | (np.abs(residual[1::2]) > temp[1::2])
        residual[np.repeat(out, 2)] = 0
        outliers_ignored += int(out.sum())
        K = P @ H.T @ np.linalg.inv(S)
        states = states + K @ residual
        est[i] = states
        P = (np.eye(n_states) - K @ H) @ P
        P_est[i] = P
    sm = est.copy()
    P_sm = P_est.copy()
    for i in range(n_frames - 2, 0, -1):
        A = P_est[i] @ F.T @ np.linalg.inv(P_pred[i + 1])
        sm[i] = est[i] + A @ (sm[i + 1] - pred[i + 1])
        P_sm[i] = P_est[i] + A @ (P_sm[i + 1] - P_pred[i + 1]) @ A.T
    return dict(x=est[:, :n], dx=est[:, n:2 * n], ddx=est[:, 2 * n:], smoothed_x=sm[:, :n], smoothed_dx=sm[:, n:2 * n],
                smoothed_ddx=sm[:, 2 * n:], outliers_ignored=outliers_ignored)
