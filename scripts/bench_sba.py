"""SBA config 4 of BASELINE.json: 6 cameras x N checkerboard views (9x6 corners), extrinsics + points.
Reports obs-evals/s of the residual+Jacobian kernel, LM iterations/s, final cost and extrinsic error."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import acinoset_b200 as ab
from acinoset_b200 import calib, sba, fte

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=5000)
args = ap.parse_args()
import torch
K, D, R, t, _ = synth.load_dummy_scene()
h = fte.get_handle(0)
p = synth.make_sba_problem(args.views, lambda X, k, d, r, tt: h.project_points(X, k, d, r, tt), seed=4)
n_pts = len(p["points_3d_true"]); n_obs = len(p["point_3d_indices"])
# initial points: two-view triangulation with the perturbed extrinsics
pts0 = np.zeros((n_pts, 3)); ci, pi = p["camera_indices"], p["point_3d_indices"]
order = np.argsort(pi, kind="stable"); first = order[np.r_[True, pi[order][1:] != pi[order][:-1]]]
second = first + 1   # observations of a point are contiguous per view/camera block ordering below
a_idx = first; b_idx = np.array([np.nonzero(pi == q)[0][1] for q in pi[first][:0]], dtype=int)
# robust pairing: for each point the first two observations in order
starts = np.nonzero(np.r_[True, pi[order][1:] != pi[order][:-1]])[0]
a_idx = order[starts]; b_idx = order[starts + 1]
for (ca, cb) in sorted(set(zip(ci[a_idx], ci[b_idx]))):
    m = (ci[a_idx] == ca) & (ci[b_idx] == cb)
    X = calib.triangulate_points_fisheye(p["points_2d"][a_idx[m]], p["points_2d"][b_idx[m]], K[ca], D[ca], p["R0"][ca], p["t0"][ca],
                                         K[cb], D[cb], p["R0"][cb], p["t0"][cb])
    pts0[pi[a_idx[m]]] = X
prob = sba.SBAProblem(p["points_2d"], pi, ci, K, D, n_pts)
x0 = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]]), p["t0"].ravel()])
# eval throughput (residual + Jacobian blocks), device resident
s = prob.st[0]
s["params"].copy_(torch.as_tensor(x0).cuda()); s["pts"].copy_(torch.as_tensor(pts0).cuda())
for _ in range(3): prob._eval(s)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): prob._eval(s)
e1.record(); torch.cuda.synchronize()
ms_eval = e0.elapsed_time(e1) / 20
prob.solve(x0, pts0, max_nfev=3, ftol=1e-10)       # warm-up: lazy kernel loading, workspace allocation
torch.cuda.synchronize()
t0 = time.perf_counter()
out = prob.solve(x0, pts0, max_nfev=200, ftol=1e-10)
dt = time.perf_counter() - t0
obj, r_new, t_new = sba.params_to_points_extrinsics(np.concatenate([out["params"], out["pts"].ravel()]), 6, n_pts)
# extrinsic error vs truth after removing the gauge (relative pose camera 0 -> camera c)
def rel(Rs, ts):
    return [(Rs[c] @ Rs[0].T, ts[c].reshape(3) - Rs[c] @ Rs[0].T @ ts[0].reshape(3)) for c in range(6)]
rt, ro, r0 = rel(p["R_true"], p["t_true"]), rel(r_new, t_new), rel(p["R0"], p["t0"])
ang = lambda A, B: float(np.degrees(np.arccos(np.clip((np.trace(A @ B.T) - 1) / 2, -1, 1))))
err_rot = max(ang(a[0], b[0]) for a, b in zip(rt, ro)); err_rot0 = max(ang(a[0], b[0]) for a, b in zip(rt, r0))
scale = np.linalg.norm(rt[2][1]) / max(np.linalg.norm(ro[2][1]), 1e-12)      # scale gauge
err_t = max(float(np.linalg.norm(a[1] - scale * b[1])) for a, b in zip(rt, ro)); err_t0 = max(float(np.linalg.norm(a[1] - b[1])) for a, b in zip(rt, r0))
print(json.dumps({"views": p["n_views"], "n_obs": n_obs, "n_pts": n_pts, "eval_ms": ms_eval, "obs_evals_per_sec": n_obs / (ms_eval * 1e-3),
                  "eval_GBps_at_176B_per_obs": n_obs * 176 / (ms_eval * 1e-3) / 1e9, "solve_s": dt, "nfev": out["nfev"], "iters": out["iters"],
                  "lm_iters_per_sec": out["nfev"] / dt, "cost0": out["cost0"], "cost": out["cost"], "status": out["status"],
                  "max_rel_rot_err_deg": [err_rot0, err_rot], "max_rel_trans_err_m": [err_t0, err_t]}))
# ---- per-kernel device times of one LM attempt (CUDA events, 20 repetitions each, state after the solve)
def timed(fn, reps=20):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps
s, t = prob.st
n6 = 6 * prob.C
lam = 1e-3
ph = {
    "sba_schur (+reduce)": lambda: prob.h.call_dev("acino_sba_schur_dev", prob.n_pts, prob.C, prob.pt_ptr, prob.obs, prob.cam_idx, s["res"], s["Jc"], s["Jp"], s["wgt"], lam, prob.partial, prob.S, prob.dc),
    "sba_dense_solve": lambda: prob.h.call_dev("acino_sba_dense_solve_dev", n6, prob.S, prob.dc, prob.info),
    "sba_backsub": lambda: prob.h.call_dev("acino_sba_backsub_dev", prob.n_pts, prob.C, prob.pt_ptr, prob.obs, prob.cam_idx, s["res"], s["Jc"], s["Jp"], s["wgt"], lam, prob.dc, s["pts"], t["pts"], prob.dp),
    "sba_pred": lambda: prob.h.call_dev("acino_sba_pred_dev", prob.n_obs, prob.cam_idx, prob.pt_idx, s["res"], s["Jc"], s["Jp"], s["wgt"], prob.dc, prob.dp, prob.pred),
    "sba_cams + sba_eval": lambda: prob._eval(t),
    "lm_reduce + norms + D2H": lambda: prob._sum(t["cost"], prob.pred, norms_of=t),
}
print(json.dumps({"phase_us": {k: round(timed(v), 1) for k, v in ph.items()}}))
