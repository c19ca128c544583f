"""LM iterations/sec of the full FTE solve (BASELINE.json configs[2]: 6 cam x 20 kpt x 10 000 frames)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import acinoset_b200 as ab
from acinoset_b200 import lm

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10000)
ap.add_argument("--init-sigma", type=float, default=0.05)
ap.add_argument("--verbose", action="store_true")
args = ap.parse_args()
import torch
K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(0); h.set_cameras(K, D, R, t)
def reproject(x):
    pos, uv = h.fk_project(x.astype(np.float32)); return pos.astype(np.float64), uv.astype(np.float64)
p = synth.make_fte_problem(args.frames, None, None, seed=3, reproject=reproject, init_sigma=args.init_sigma)
sol = lm.FTESolver(h, p["meas"], p["w"], p["Ts"])
x, info = sol.solve(p["x0"], max_iter=3)   # warm-up (module load, allocations)
torch.cuda.synchronize()
t0 = time.perf_counter()
x, info = sol.solve(p["x0"], max_iter=60, verbose=args.verbose)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pos, _ = h.fk_project(x.astype(np.float32)); post, _ = h.fk_project(p["x_true"].astype(np.float32))
rms = float(np.sqrt(((pos - post) ** 2).sum(-1).mean()))
print(json.dumps({"frames": args.frames, "iters": info["iters"], "attempts": info["n_solve"], "evals": info["n_eval"],
                  "seconds": dt, "lm_iters_per_sec": info["n_solve"] / dt, "ms_per_attempt": 1e3 * dt / info["n_solve"],
                  "F": info["F"], "converged": info["converged"], "marker_rms_m": rms, "launches": info["launches"]}))
