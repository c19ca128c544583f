"""Independent-optimiser pins for the two solves: SciPy drives the ORACLE's objective /
residuals with the oracle's analytic derivatives - a different algorithm (More-Sorensen trust-region Newton / trust-region
reflective with an exact SVD sub-problem) from the GPU's projected Levenberg-Marquardt - and the end points are committed
as tests/golden/solves.npz.  Runs in the build container (CPU, SciPy); needs neither /root/reference nor a GPU:

    python tests/golden/make_golden_solves.py            # ~1 minute

* FTE (replaces Pyomo + IPOPT, all_optimizations.py:503-524, absent from the image): scipy.optimize.minimize(
  method="trust-exact") on oracle.fte.total_objective with the exact gradient (oracle fte_eval g + smooth_grad) and the
  Gauss-Newton matrix B = blockdiag(H_n) + S as `hess`, from the same initial guess synth.make_fte_problem gives the GPU
  (seeded), N = 48 and 100 frames.  trust-exact takes no bounds: the 21 bounds of :403-483 are checked to be INACTIVE at
  the end point, so it is a stationary point of the bound-constrained problem too.  (minimize(method="trust-constr") WITH
  the bounds - an interior-point method - was run as well: after 700 iterations / 11 minutes at N = 48 it had reached
  F = 19913.104 and was still descending towards the 19912.9968 found here in 25 iterations; too slow to be the pin.)
* SBA K6 / K7 (calib.py:369-390): scipy.optimize.least_squares with the REFERENCE's options (method='trf', loss='cauchy',
  x_scale='jac', ftol=1e-10, max_nfev=1000) but jac = the oracle's analytic Jacobian instead of the reference's
  finite differences over a sparsity pattern whose columns do not match its parameter layout (calib.py:202 vs :346-351),
  and tr_solver='exact' (dense SVD): with the default LSMR sub-problem solver the gauge freedom (7 unobservable
  directions) makes TRF crawl - 15 000 evaluations to get within 2.5e-3 of the cost the exact solver reaches in 28.
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.optimize import least_squares, minimize

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from oracle import fisheye, fte, lm as olm, sba as osba, skeleton  # noqa: E402

NA = 25


def fte_pin(N, seed):
    cams = synth.load_dummy_scene()
    K, D, R, t, _ = cams
    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=seed, cams=cams)
    q = fte.model_weights_active()
    lo, hi = skeleton.active_bounds()
    S = olm.smooth_matrix(N, p["Ts"], q).toarray()
    # what the GPU evaluates: measurements / weights rounded to fp32 (the state stays fp64 here)
    meas = p["meas"].astype(np.float32).astype(np.float64)
    w = p["w"].astype(np.float32).astype(np.float64)
    cache = {}

    def ev(xv):
        key = xv.tobytes()
        if key not in cache:
            cache.clear()
            x = xv.reshape(N, NA)
            c, g, H = fte.fte_eval(x, meas, w, K, D, R, t)
            cache[key] = (float(c.sum()) + fte.smooth_cost(x, p["Ts"], q), (g + fte.smooth_grad(x, p["Ts"], q)).ravel(), H)
        return cache[key]

    def hess(xv):
        B = S.copy()
        for n, Hn in enumerate(ev(xv)[2]):
            B[n * NA:(n + 1) * NA, n * NA:(n + 1) * NA] += Hn
        return B

    x0 = np.clip(p["x0"], lo, hi).ravel()
    t0 = time.time()
    res = minimize(lambda v: ev(v)[0], x0, jac=lambda v: ev(v)[1], hess=hess, method="trust-exact",
                   options=dict(gtol=1e-7, maxiter=500))
    x = res.x.reshape(N, NA)
    inside = bool(np.all((x > lo + 1e-9) & (x < hi - 1e-9)))
    F, g = ev(res.x)[0], ev(res.x)[1]
    print(f"FTE N={N}: F0 {ev(x0)[0]:.6f} -> F {F:.8f}, {res.nit} iterations, {time.time() - t0:.1f} s, "
          f"|gradient|inf {np.abs(g).max():.3e}, all bounds inactive: {inside}")
    assert inside, "a bound is active at the trust-exact end point: it is not a pin for the bounded problem"
    return dict(x=x, F=F, seed=seed, pg=np.abs(g).max())


def sba_pin(g, tag):
    K, D = g[f"{tag}_K"], g[f"{tag}_D"].reshape(-1, 4)
    pidx, cidx, p2d = g[f"{tag}_pidx"], g[f"{tag}_cidx"], g[f"{tag}_points_2d"].astype(np.float64)
    n_pts = len(g[f"{tag}_points_3d"])
    x0 = g[f"{tag}_x0"]
    n_obs = len(pidx)
    rows = np.repeat(np.arange(2 * n_obs), 9).reshape(n_obs, 2, 9)
    cols = np.empty((n_obs, 2, 9), dtype=np.int64)
    for k in range(3):
        cols[:, :, k] = (3 * cidx + k)[:, None]
        cols[:, :, 3 + k] = (6 + 3 * cidx + k)[:, None]
        cols[:, :, 6 + k] = (12 + 3 * pidx + k)[:, None]

    def fun(v):
        return osba.cost_func_points_extrinsics(v, 2, n_pts, pidx, cidx, K, D, p2d)

    def jac(v):
        Jr, Jt, Jx = osba.jac_blocks_points_extrinsics(v, 2, n_pts, pidx, cidx, K, D)
        vals = np.concatenate([Jr, Jt, Jx], axis=2)
        return sp.csr_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(2 * n_obs, 12 + 3 * n_pts))

    t0 = time.time()
    res = least_squares(fun, x0, jac=lambda v: jac(v).toarray(), x_scale="jac", ftol=1e-10, method="trf", loss="cauchy",
                        max_nfev=1000, tr_solver="exact", verbose=0)
    cost = 0.5 * np.sum(np.log1p(res.fun ** 2))
    print(f"SBA {tag}: cost {0.5 * np.sum(np.log1p(fun(x0) ** 2)):.6e} -> {cost:.10e}, nfev {res.nfev}, status {res.status}, "
          f"optimality {res.optimality:.3e}, {time.time() - t0:.1f} s")
    # the reference's pattern bug, quantified: true non-zeros of the analytic Jacobian outside its sparsity pattern
    A = osba.sparsity(2, 6, cidx, n_pts, pidx)
    J0 = jac(x0).toarray()
    outside = int(np.count_nonzero((np.abs(J0) > 1e-12) & (A == 0)))
    print(f"   analytic non-zeros outside the reference's pattern at x0: {outside} of {int(np.count_nonzero(np.abs(J0) > 1e-12))}")
    return dict(x=res.x, cost=cost, nfev=res.nfev, optimality=res.optimality, outside=outside)


def main():
    out = {}
    for N, seed in ((48, 5), (100, 6)):
        r = fte_pin(N, seed)
        out[f"fte{N}_x"], out[f"fte{N}_F"], out[f"fte{N}_seed"], out[f"fte{N}_pg"] = r["x"], r["F"], r["seed"], r["pg"]
    g = np.load(os.path.join(HERE, "sba.npz"))
    for tag in ("static", "rotating"):
        r = sba_pin(g, tag)
        out[f"sba_{tag}_x"], out[f"sba_{tag}_cost"] = r["x"], r["cost"]
        out[f"sba_{tag}_nfev"], out[f"sba_{tag}_optimality"], out[f"sba_{tag}_outside"] = r["nfev"], r["optimality"], r["outside"]
    np.savez_compressed(os.path.join(HERE, "solves.npz"), **out)
    print("wrote", os.path.join(HERE, "solves.npz"))


if __name__ == "__main__":
    main()
