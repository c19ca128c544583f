"""Oracle (test infrastructure): fp64 CPU restatement of the Levenberg-Marquardt loop that
replaces the reference's Pyomo + IPOPT solve (/root/reference/src/all_optimizations.py:503-524)
on the reduced FTE objective of SURVEY.md appendix B6:

    F(x) = sum_{n,c,l,d} rho(w (proj - meas)) + sum_{n>=3,p} q_p (third difference / Ts^2)^2
    subject to the 21 box bounds of all_optimizations.py:403-483.

PARITY UNPINNED for the solve itself: pyomo/ipopt are not installable in the build image, so
there is no reference solution to pin against; the objective / gradient it consumes are pinned
(tests/test_oracle_golden.py).  The CUDA solver (acinoset_b200.fte.fte_solve) runs the same
algorithm and is compared with this restatement.

Algorithm (identical in the CUDA path): damped Gauss-Newton on B = blockdiag(H_n) + S with
Marquardt scaling; (B + lam diag(B)) dx = -g; bound-fixed variables (at a bound with the
gradient pushing outwards) are frozen; trial = clip(x + dx); gain ratio against the model;
lam /= 3 on success (>= 0.75: /= 3, < 0.25: *= 2), lam *= 4 on rejection.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fte, skeleton

NA = skeleton.N_ACTIVE


def smooth_matrix(N, Ts, q):
    """Sparse Hessian of the smoothness term: 2 q_p/Ts^4 D3^T D3 (x) per parameter."""
    if N < 4:
        return sp.csr_matrix((N * NA, N * NA))
    rows, cols, vals = [], [], []
    st = np.array([-1.0, 3.0, -3.0, 1.0])
    for k in range(4):
        rows.append(np.arange(N - 3))
        cols.append(np.arange(N - 3) + k)
        vals.append(np.full(N - 3, st[k]))
    D3 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N - 3, N))
    G = (D3.T @ D3).tocsr()
    return sp.kron(G, sp.diags(2 * q / Ts ** 4), format="csr")


def assemble(H, N):
    return sp.block_diag([sp.csr_matrix(H[n]) for n in range(N)], format="csr")


def objective(xa, prob, q):
    K, D, R, t, _ = prob["cams"]
    c, g, H = fte.fte_eval(xa, prob["meas"], prob["w"], K, D, R, t)
    return float(c.sum()) + fte.smooth_cost(xa, prob["Ts"], q), c, g, H


def solve(prob, x0, max_iter=60, lam0=1e-3, tol_step=1e-6, tol_rel=1e-8, q=None, verbose=False,
          eval_dtype=np.float64):
    """Returns (x, info).  prob: dict from synth.make_fte_problem (meas, w, cams, Ts).

    eval_dtype=np.float32 rounds the state / measurements to fp32 before every evaluation
    (what the CUDA path does) while the iterate itself stays fp64."""
    q = fte.model_weights_active() if q is None else q
    lo, hi = skeleton.active_bounds()
    K, D, R, t, _ = prob["cams"]
    N = x0.shape[0]
    Ts = prob["Ts"]
    S = smooth_matrix(N, Ts, q)
    meas = prob["meas"].astype(eval_dtype).astype(np.float64)
    w = prob["w"].astype(eval_dtype).astype(np.float64)

    def ev(x):
        xe = x.astype(eval_dtype).astype(np.float64)
        c, g, H = fte.fte_eval(xe, meas, w, K, D, R, t)
        return float(c.sum()) + fte.smooth_cost(x, Ts, q), g + fte.smooth_grad(x, Ts, q), H

    x = np.clip(np.asarray(x0, dtype=np.float64), lo, hi)
    F, g, H = ev(x)
    lam = lam0
    hist = [F]
    n_eval = 1
    for it in range(max_iter):
        B = (assemble(H, N) + S).tocsr()
        gv = g.ravel().copy()
        xv = x.ravel()
        lov, hiv = np.tile(lo, N), np.tile(hi, N)
        fixed = ((xv <= lov) & (gv > 0)) | ((xv >= hiv) & (gv < 0))
        dB = B.diagonal()
        accepted = False
        for _ in range(12):
            A = (B + sp.diags(lam * dB)).tolil()
            # freeze bound-fixed variables
            idx = np.nonzero(fixed)[0]
            A = A.tocsr()
            if idx.size:
                keep = np.ones(N * NA)
                keep[idx] = 0
                Dk = sp.diags(keep)
                A = Dk @ A @ Dk + sp.diags(1 - keep)
            rhs = -gv.copy()
            rhs[fixed] = 0
            dx = spla.spsolve(A.tocsc(), rhs)
            xt = np.clip(xv + dx, lov, hiv)
            dxe = xt - xv
            pred = -(gv @ dxe) - 0.5 * dxe @ (B @ dxe)
            Ft, gt, Ht = ev(xt.reshape(N, NA))
            n_eval += 1
            rho = (F - Ft) / pred if pred > 0 else -1.0
            if verbose:
                print(f"it {it:3d} lam {lam:9.3e} F {F:14.6f} Ft {Ft:14.6f} pred {pred:10.3e} rho {rho:6.3f} |dx|inf {np.abs(dxe).max():.2e}")
            if Ft < F and rho > 1e-4:
                accepted = True
                step = np.abs(dxe).max()
                rel = (F - Ft) / max(abs(F), 1e-30)
                x, F, g, H = xt.reshape(N, NA), Ft, gt, Ht
                if rho > 0.75:
                    lam = max(lam / 3, 1e-12)
                elif rho < 0.25:
                    lam = lam * 2
                break
            lam *= 4
        hist.append(F)
        if not accepted:
            break
        if step < tol_step or rel < tol_rel:
            break
    return x, dict(F=F, iters=it + 1, n_eval=n_eval, history=hist, lam=lam)
