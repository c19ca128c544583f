"""Round-2 kernels under compute-sanitizer: the persistent fte_eval with several tiles per CTA (+ a partial last tile, + the
unaligned / non-bulk path) and the SBA Schur kernel with its shared-memory transposition.
    compute-sanitizer --tool memcheck  python scripts/sanitize_r02.py
    compute-sanitizer --tool racecheck python scripts/sanitize_r02.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import synth
import acinoset_b200 as ab
from acinoset_b200 import sba
from oracle import fisheye

K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(0); h.set_cameras(K, D, R, t)
n = 8 * 592 * 2 + 8 * 37 + 3                      # every CTA walks >= 2 tiles, some 3; the last tile is partial
rng = np.random.default_rng(0)
x = synth.make_trajectory(n, rng).astype(np.float32)
pos, uv = h.fk_project(x)
meas = (uv + rng.normal(0, 2, uv.shape)).astype(np.float32); w = np.full((n, 6, 20), 0.2, np.float32)
w[rng.random(w.shape) < 0.1] = 0.0
c1, g1, H1 = h.fte_eval(x, meas, w)
# same through device tensors whose base is NOT 16-byte aligned: the plain-load / plain-store path of every tile
dev = torch.device("cuda:0")
def off(a):
    buf = torch.empty(a.size + 1, dtype=torch.float32, device=dev)
    v = buf[1:].view(a.shape); v.copy_(torch.from_numpy(a)); return v
xd, wd = off(x), off(w)
md = torch.from_numpy(meas).to(dev)               # meas must stay 8-byte aligned (float2 loads)
cd, gd, Hd = off(np.zeros(n, np.float32)), off(np.zeros((n, 25), np.float32)), off(np.zeros((n, 325), np.float32))
h.fte_eval_dev(xd, md, wd, cd, gd, Hd)
torch.cuda.synchronize()
print("fte_eval", n, "frames; staged vs plain path max |dH| rel", float(np.abs(Hd.cpu().numpy() - H1).max() / np.abs(H1).max()),
      "cost", float(c1.sum()), float(cd.sum().item()))
cn, gn, _ = h.fte_eval(x, meas, w, want_H=False)
print("no-H variant: |dg|", float(np.abs(gn - g1).max()))
p = synth.make_sba_problem(40, fisheye.project, seed=2)
n_pts = len(p["points_3d_true"])
prob = sba.SBAProblem(p["points_2d"], p["point_3d_indices"], p["camera_indices"], p["K"], p["D"], n_pts)
x0 = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]]), p["t0"].ravel()])
out = prob.solve(x0, p["points_3d_true"] + 0.01, max_nfev=4)
print("sba", out["cost0"], "->", out["cost"], out["nfev"])
