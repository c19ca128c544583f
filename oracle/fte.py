"""Oracle (test infrastructure): the reduced FTE objective, its gradient and its
Gauss-Newton blocks, fp64 NumPy.

Follows /root/reference/src/all_optimizations.py
  * pose_constraint            :359-365   poses[n,l,:] = FK_l(x[n,:])
  * measurement_constraints    :394-399   slack_meas = proj_d(pose; cam c) - meas
  * init_meas_weights          :302-308   w = 1/R (R=5) if likelihood > dlc_thresh else 0
  * backwards_euler_pos / _vel / constant_acc :369-391  => slack_model[n] = third
                                          backward difference of x / Ts^2 (n >= 4, 1-based)
  * init_model_weights, Q      :245-252,310-315
  * obj                        :486-500   sum w_p slack_model^2 + sum redesc(w * slack_meas)
Per-frame output of ``fte_eval``: cost_n, g_n = d cost_n / d x_n (25), and
H_n = sum psi(e) w^2 J^T J (25x25, psi = curvature weight of oracle/loss.py).
"""
import numpy as np

from . import fisheye, loss, skeleton

MEAS_SIGMA_R = 5.0  # all_optimizations.py:243
NA = skeleton.N_ACTIVE
N_UPPER = NA * (NA + 1) // 2  # 325


def upper_index(i, j):
    """Packed row-major upper-triangular index of (i, j), i <= j, for a 25x25 block."""
    return i * NA - (i * (i - 1)) // 2 + (j - i)


def pack_upper(H):
    iu = np.triu_indices(NA)
    return np.asarray(H)[..., iu[0], iu[1]]


def unpack_upper(Hu):
    Hu = np.asarray(Hu)
    iu = np.triu_indices(NA)
    H = np.zeros(Hu.shape[:-1] + (NA, NA), dtype=Hu.dtype)
    H[..., iu[0], iu[1]] = Hu
    H[..., iu[1], iu[0]] = Hu
    return H


def meas_weights(likelihood, dlc_thresh, sigma=MEAS_SIGMA_R):
    """init_meas_weights: 1/R where likelihood > thresh else 0."""
    return np.where(np.asarray(likelihood) > dlc_thresh, 1.0 / sigma, 0.0)


def model_weights_active():
    """1/Q_p for the 25 active slots (Q = sigma^2), all_optimizations.py:245-252,310-315."""
    s = skeleton.Q_SIGMA[skeleton.ACTIVE_IDX]
    return 1.0 / (s * s)


def reproject(xa, K, D, R, t):
    """(N,25) -> pixels (N,C,L,2)."""
    P = skeleton.cheetah_fk_active(xa)
    C = len(K)
    return np.stack([fisheye.project(P, K[c], D[c], R[c], t[c]) for c in range(C)], axis=-3)


def residuals_and_jac(xa, meas, K, D, R, t):
    """r (N,C,L,2) = proj - meas ;  J (N,C,L,2,25) = d r / d x_active."""
    xa = np.asarray(xa, dtype=np.float64)
    P = skeleton.cheetah_fk_active(xa)
    Jfk = skeleton.cheetah_fk_jac(xa)  # (N,L,3,25)
    C = len(K)
    rs, Js = [], []
    for c in range(C):
        uv, Jw, _ = fisheye.project_jac(P, K[c], D[c], R[c], t[c])  # (N,L,2), (N,L,2,3)
        rs.append(uv - meas[:, c])
        Js.append(Jw @ Jfk)
    return np.stack(rs, axis=1), np.stack(Js, axis=1)


def fte_eval(xa, meas, w, K, D, R, t, abc=(loss.REDESC_A, loss.REDESC_B, loss.REDESC_C)):
    """Per-frame measurement cost, gradient and GN block.

    xa (N,25) f64, meas (N,C,L,2), w (N,C,L).  Returns cost (N,), g (N,25), H (N,25,25).
    """
    r, J = residuals_and_jac(xa, meas, K, D, R, t)
    e = w[..., None] * r  # (N,C,L,2)
    rho, drho, _ = loss.redescending_dloss(e, *abc)
    psi = loss.redescending_irls_weight(e, *abc)
    cost = rho.sum(axis=(1, 2, 3))
    gam = drho * w[..., None]            # d rho / d r
    eta = psi * (w * w)[..., None]
    g = np.einsum("ncld,ncldp->np", gam, J)
    H = np.einsum("ncld,ncldp,ncldq->npq", eta, J, J)
    return cost, g, H


def smooth_cost(xa, Ts, q=None):
    """sum_{n>=3} sum_p q_p (third backward difference / Ts^2)^2  (0-based n)."""
    q = model_weights_active() if q is None else q
    xa = np.asarray(xa, dtype=np.float64)
    if xa.shape[0] < 4:
        return 0.0
    d3 = (xa[3:] - 3 * xa[2:-1] + 3 * xa[1:-2] - xa[:-3]) / (Ts * Ts)
    return float(np.sum(q * d3 * d3))


def smooth_grad(xa, Ts, q=None):
    q = model_weights_active() if q is None else q
    xa = np.asarray(xa, dtype=np.float64)
    g = np.zeros_like(xa)
    if xa.shape[0] < 4:
        return g
    d3 = (xa[3:] - 3 * xa[2:-1] + 3 * xa[1:-2] - xa[:-3]) * (2 * q / Ts ** 4)
    g[3:] += d3
    g[2:-1] += -3 * d3
    g[1:-2] += 3 * d3
    g[:-3] += -d3
    return g


def smooth_band(N, Ts, q=None):
    """Heptadiagonal bands of 2 q_p/Ts^4 * D3^T D3: array (N, 4, 25) with
    band[n,k,p] = Hessian entry between frames n and n+k for parameter p."""
    q = model_weights_active() if q is None else q
    band = np.zeros((N, 4))
    stencil = np.array([-1.0, 3.0, -3.0, 1.0])
    for n in range(3, N):
        rows = [n - 3, n - 2, n - 1, n]
        for i in range(4):
            for j in range(i, 4):
                band[rows[i], j - i] += stencil[i] * stencil[j]
    return band[:, :, None] * (2 * q / Ts ** 4)


def total_objective(xa, meas, w, K, D, R, t, Ts, q=None):
    cost, _, _ = fte_eval(xa, meas, w, K, D, R, t)
    return float(cost.sum()) + smooth_cost(xa, Ts, q)


def derived_velocities(xa, Ts):
    """dx, ddx of the output pickle."""
    xa = np.asarray(xa, dtype=np.float64)
    N = xa.shape[0]
    dx = np.zeros_like(xa)
    ddx = np.zeros_like(xa)
    dx[1:] = (xa[1:] - xa[:-1]) / Ts
    if N > 2:
        ddx[2:] = (dx[2:] - dx[1:-1]) / Ts
        ddx[1] = ddx[2]
        ddx[0] = ddx[1]
    dx[0] = dx[1] - Ts * ddx[1] if N > 1 else 0.0
    return dx, ddx
