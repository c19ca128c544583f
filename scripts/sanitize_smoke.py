"""Tiny end-to-end pass over the generic-skeleton and stereo kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth
from acinoset_b200 import build, fte, skeleton, stereo

g = np.load(os.path.join(ROOT, "tests", "golden", "generic_fk.npz"))
skel = json.loads(str(g["K1_skeleton_json"]))
flat = skeleton.flatten_skeleton(skel)
K, D, R, t, _ = synth.load_dummy_scene()
K, D, R, t = K[:3], D[:3], R[:3], t[:3]
rng = np.random.default_rng(0)
N = 6
x = np.array(g["K1_x"][:N], dtype=np.float64)
x[:, :3] = [2.0, 6.5, 1.0]
h = fte.set_scene(K, D, R, t)
pose = skeleton.build_pose_function(skel)
P3 = pose(x)
uv = np.stack([h.project_points(P3.reshape(-1, 3), K[c], D[c], R[c], t[c]).reshape(N, -1, 2) for c in range(3)], 1)
meas = uv + rng.normal(0, 1.0, uv.shape)
w = np.full(meas.shape[:-1], 1 / 3.0)
solver = build.SkelSolver(h, flat, meas, w, 1 / 120.0, loss="abs")
xs, info = solver.solve(x + rng.normal(0, 0.02, x.shape), max_iter=2)
print("skel", info["F0"], info["F"], info["n_solve"])
s = np.load(os.path.join(ROOT, "tests", "golden", "stereo.npz"))
rms, Rr, Tr = stereo.solve_pair(s["rot12_obj"], s["rot12_img1"][:4], s["rot12_img2"][:4], s["rot12_K1"], s["rot12_D1"], s["rot12_K2"],
                                s["rot12_D2"], max_iter=3)
print("stereo", rms)
