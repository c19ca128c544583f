"""Small pass over the cheetah kernels (fte_eval, fte_jac, fk_project, LM + BCR, SBA, triangulation) for compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_core.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth
import acinoset_b200 as ab
from acinoset_b200 import calib, lm

K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(0); h.set_cameras(K, D, R, t)
def reproject(x):
    pos, uv = h.fk_project(x.astype(np.float32)); return pos.astype(np.float64), uv.astype(np.float64)
p = synth.make_fte_problem(96, None, None, seed=3, reproject=reproject)
cost, g, H = h.fte_eval(p["x0"].astype(np.float32), p["meas"].astype(np.float32), p["w"].astype(np.float32))
uv, J = h.fte_jac(p["x0"][:16].astype(np.float32))
print("fte_eval", float(cost.sum()), "fte_jac", J.shape)
sol = lm.FTESolver(h, p["meas"], p["w"], p["Ts"])
x, info = sol.solve(p["x0"], max_iter=3)
print("lm", info["F"], info["n_solve"], info["bcr_info"])
valid = p["lik"] > 0.5
pos, cnt = calib.triangulate_pairwise_dense(p["meas"].astype(np.float64), valid, K, D, R, t)
print("tri", int(cnt.sum()))
