"""Oracle (test infrastructure): the generic-skeleton FTE variant of the reference
(/root/reference/src/build.py:28-302), fp64 NumPy.

Follows
  * build.py:43-80     forward kinematics from a skeleton pickle (quirks kept: local rotations do not chain,
                       ``rot[child + "_i"]`` toggles each time a part appears as a child, a child listed twice
                       is overwritten) - evaluated here with complex arithmetic so that
  * the Jacobian of every residual is obtained by COMPLEX-STEP differentiation (h = 1e-30: exact to rounding),
    a route independent of the kernel's analytic link formulation (csrc/skel_body.cuh);
  * build.py:457-473   pt3d_to_2d;  :166-172 measurement weights 1/R (R = 3) above the 0.4 likelihood threshold,
                       marker "neck" skipped (:123-124,196-197,281-282);
  * build.py:287-302   objective  sum 0.002 slack_model^2 + sum |w slack_meas|   (``loss="abs"``), or the
                       redescending loss of all_optimizations.py:497 (``loss="redescending"``);
  * build.py:231-261   backwards_euler_pos / _vel + constant_acc  =>  slack_model = third difference / h^2
                       (SURVEY.md appendix B6);  :263-266 bounds |x_i| <= pi/2 for 1-based i in [3, 3L), frames 1..N-1.
Pinning: the FK is pinned to the reference's shipped result pickles (K1/K2, tests/golden/generic_fk.npz); the
objective pieces reuse the pinned projection and loss.  The IPOPT solve itself is PARITY UNPINNED (pyomo / ipopt
absent): `solve` below is an fp64 CPU run of the same Levenberg-Marquardt algorithm as the CUDA path.
"""
import numpy as np

from . import loss as _loss

MODEL_WEIGHT = 0.002       # build.py:176
MEAS_SIGMA_R = 3.0         # build.py:142
LIK_THRESH = 0.4           # build.py:166-170


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])      # build.py:399-405
    if axis == 1:
        return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])      # :407-414
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])          # :416-423


def pose_function(skel_dict):
    """-> (pose_to_3d(x) -> (n_out, 3), out_names); x = [x,y,z,*phi,*theta,*psi] real or complex."""
    links = skel_dict["links"]
    positions = skel_dict["positions"]
    dofs = {k: list(v) for k, v in skel_dict["dofs"].items()}
    for joint in skel_dict["markers"]:           # build.py:37-38
        dofs[joint] = [1, 1, 1]
    parts = list(dofs.keys())
    L = len(positions)
    names = []
    for link in links:
        for p in link:
            if p not in names:
                names.append(p)

    def pose_to_3d(x):
        x = np.asarray(x)
        rot_i = {}
        for i, part in enumerate(parts):         # build.py:51-62
            Rm = np.eye(3, dtype=x.dtype)
            if dofs[part][1]:
                Rm = _rot(1, x[3 + L + i]) @ Rm
            if dofs[part][0]:
                Rm = _rot(0, x[3 + i]) @ Rm
            if dofs[part][2]:
                Rm = _rot(2, x[3 + 2 * L + i]) @ Rm
            rot_i[part] = Rm.T
        pose = {}
        for link in links:                       # build.py:68-80
            if len(link) == 1:
                pose[link[0]] = x[0:3].copy()
                continue
            a, b = link
            if a not in pose:
                pose[a] = x[0:3].copy()
            tv = np.asarray(positions[b], dtype=np.float64) - np.asarray(positions[a], dtype=np.float64)
            rot_i[b] = rot_i[b].T
            pose[b] = pose[a] + rot_i[a] @ tv
        return np.stack([pose[k] for k in pose], axis=0)

    return pose_to_3d, names


def project(X, K, D, R, t):
    """pt3d_to_2d (build.py:457-473), complex-capable.  X (..., 3) -> (..., 2)."""
    Xc = X @ np.asarray(R).T + np.asarray(t).reshape(3)
    a = Xc[..., 0] / Xc[..., 2]
    b = Xc[..., 1] / Xc[..., 2]
    r = np.sqrt(a * a + b * b + 1e-12)
    th = np.arctan(r)
    D = np.asarray(D).reshape(4)
    thd = th * (1 + D[0] * th ** 2 + D[1] * th ** 4 + D[2] * th ** 6 + D[3] * th ** 8)
    return np.stack([K[0][0] * a * thd / r + K[0][2], K[1][1] * b * thd / r + K[1][2]], axis=-1)


def residuals(skel_dict, x, meas, K, D, R, t):
    """x (P,), meas (C, n_out, 2) -> r (C, n_out, 2) (projection - measurement)."""
    f, _ = pose_function(skel_dict)
    P3 = f(x)
    return np.stack([project(P3, K[c], D[c], R[c], t[c]) for c in range(len(K))], 0) - meas


def residual_jacobian(skel_dict, x, meas, K, D, R, t, h=1e-30):
    """Complex-step Jacobian of the residuals: (C, n_out, 2, P)."""
    x = np.asarray(x, dtype=np.float64)
    J = np.empty(meas.shape + (x.size,))
    for p in range(x.size):
        xc = x.astype(np.complex128)
        xc[p] += 1j * h
        J[..., p] = residuals(skel_dict, xc, meas, K, D, R, t).imag / h
    return J


def loss_terms(e, kind, abc, delta):
    """e = |w r| >= 0 -> (rho, gw = rho'/e, hw = curvature weight)."""
    e = np.asarray(e, dtype=np.float64)
    pos = e > 0
    if kind == "abs":
        rho = e.copy()
        gw = np.where(pos, 1.0 / np.where(pos, e, 1.0), 0.0)
        hw = 1.0 / np.maximum(e, delta)
        return rho, gw, hw
    a, b, c = abc
    rho = _loss.redescending_loss(e, a, b, c)
    _, d, _ = _loss.redescending_dloss(e, a, b, c)
    fl = 1.0 - _loss.func_step(a, e)
    gw = np.where(pos, d / np.where(pos, e, 1.0), 0.0)
    hw = np.where(pos, np.maximum(gw, fl), fl)
    return rho, gw, hw


def upper_pack(H):
    P = H.shape[-1]
    iu = np.triu_indices(P)
    return H[..., iu[0], iu[1]]


def upper_unpack(Hu, P):
    H = np.zeros(Hu.shape[:-1] + (P, P))
    iu = np.triu_indices(P)
    H[..., iu[0], iu[1]] = Hu
    H[..., iu[1], iu[0]] = Hu
    return H


def skel_eval(skel_dict, x, meas, w, K, D, R, t, loss="abs", abc=(3.0, 10.0, 20.0), delta=0.05):
    """x (N,P), meas (N,C,n_out,2), w (N,C,n_out) -> cost (N,), g (N,P), H (N,P,P) of the measurement term."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    N, P = x.shape
    cost = np.zeros(N)
    g = np.zeros((N, P))
    H = np.zeros((N, P, P))
    for n in range(N):
        r = residuals(skel_dict, x[n], meas[n], K, D, R, t)
        J = residual_jacobian(skel_dict, x[n], meas[n], K, D, R, t)
        wn = w[n][..., None]
        r = np.where(wn != 0, r, 0.0)
        rho, gw, hw = loss_terms(np.abs(wn * r), loss, abc, delta)
        cost[n] = rho.sum()
        g[n] = np.einsum("cld,cldp->p", gw * wn ** 2 * r, J)
        H[n] = np.einsum("cld,cldp,cldq->pq", hw * wn ** 2 * np.ones_like(r), J, J)
    return cost, g, H


def bounds(n_parts):
    """build.py:263-266: |x_i| <= pi/2 for 1-based i in [3, 3L) - which includes z (i = 3)."""
    P = 3 + 3 * n_parts
    lo = np.full(P, -np.inf)
    hi = np.full(P, np.inf)
    lo[2:3 * n_parts - 1] = -np.pi / 2
    hi[2:3 * n_parts - 1] = np.pi / 2
    return lo, hi


def d3_matrix(N):
    D3 = np.zeros((max(N - 3, 0), N))
    for m in range(3, N):
        D3[m - 3, m - 3:m + 1] = [-1.0, 3.0, -3.0, 1.0]
    return D3


def smooth_cost(x, sw):
    """0.5 sw (third difference)^2 summed; sw = 2 q / h^4."""
    D3 = d3_matrix(x.shape[0])
    return float(0.5 * np.sum(sw * (D3 @ x) ** 2))


def solve(skel_dict, x0, meas, w, cams, h, loss="abs", abc=(3.0, 10.0, 20.0), delta=0.05, max_iter=50, lam0=1e-3,
          tol_step=1e-6, tol_rel=1e-8, last_free=True, verbose=False):
    """fp64 CPU run of the LM algorithm of the CUDA path (acinoset_b200.build.solve_optimisation)."""
    K, D, R, t = cams
    x = np.array(x0, dtype=np.float64)
    N, P = x.shape
    L = (P - 3) // 3
    sw = np.full(P, 2 * MODEL_WEIGHT / h ** 4)
    lo, hi = bounds(L)
    lo_f = np.tile(lo, (N, 1))
    hi_f = np.tile(hi, (N, 1))
    if last_free:
        lo_f[-1] = -np.inf
        hi_f[-1] = np.inf
    x = np.clip(x, lo_f, hi_f)
    G3 = d3_matrix(N).T @ d3_matrix(N)
    S = np.kron(G3, np.diag(sw))

    def objective(xx):
        c, g, H = skel_eval(skel_dict, xx, meas, w, K, D, R, t, loss, abc, delta)
        return c.sum() + smooth_cost(xx, sw), g, H

    F, g, H = objective(x)
    lam = lam0
    info = dict(iters=0, attempts=0, converged=False, F0=F)
    for it in range(max_iter):
        gtot = g + (G3 @ x) * sw
        fixed = ((x <= lo_f) & (gtot > 0)) | ((x >= hi_f) & (gtot < 0))
        B = S.copy()
        for n in range(N):
            B[n * P:(n + 1) * P, n * P:(n + 1) * P] += H[n]
        accepted = False
        for _ in range(12):
            info["attempts"] += 1
            Bd = B + lam * np.diag(np.diag(B))
            fx = fixed.ravel()
            Bd[fx, :] = 0
            Bd[:, fx] = 0
            Bd[fx, fx] = 1
            rhs = np.where(fx, 0.0, -gtot.ravel())
            d = np.linalg.solve(Bd, rhs).reshape(N, P)
            xt = np.clip(x + d, lo_f, hi_f)
            s = xt - x
            pred = -(gtot * s).sum() - 0.5 * s.ravel() @ (B @ s.ravel())
            Ft, gt, Ht = objective(xt)
            rho = (F - Ft) / pred if pred > 0 else -1.0
            if verbose:
                print(f"it {it} lam {lam:.2e} F {F:.6f} Ft {Ft:.6f} pred {pred:.3e} rho {rho:.3f}")
            if Ft < F and rho > 1e-4:
                step = np.abs(s).max()
                rel = (F - Ft) / max(abs(F), 1e-300)
                x, F, g, H = xt, Ft, gt, Ht
                lam = lam / 3 if rho > 0.75 else (lam * 2 if rho < 0.25 else lam)
                accepted = True
                break
            lam *= 4
        info["iters"] = it + 1
        if not accepted:
            break
        if step < tol_step or rel < tol_rel:
            info["converged"] = True
            break
    info["F"] = F
    return x, info
