"""Generic-skeleton variant (reference src/build.py: 100 frames x 48 parameters, 4 cameras): evaluation rate and
LM solve time on one GPU.  python scripts/bench_skel.py [--frames 100] [--eval-frames 20000]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--eval-frames", type=int, default=20000)
args = ap.parse_args()
import synth
from acinoset_b200 import build, fte, skeleton

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "generic_fk.npz"))
skel = json.loads(str(g["K1_skeleton_json"]))
flat = skeleton.flatten_skeleton(skel)
K, D, R, t, _ = synth.load_dummy_scene()
K, D, R, t = K[:4], D[:4], R[:4], t[:4]
rng = np.random.default_rng(0)
N = args.frames
x_true = np.array(g["K1_x"][:N], dtype=np.float64)
tt = np.arange(N) / 120.0
x_true[:, 0] = 2.0 + 1.5 * tt; x_true[:, 1] = 6.5 + 0.8 * tt; x_true[:, 2] = 1.0
pose = skeleton.build_pose_function(skel)
P3 = pose(x_true)
h = fte.set_scene(K, D, R, t)
uv = np.stack([h.project_points(P3.reshape(-1, 3), K[c], D[c], R[c], t[c]).reshape(N, -1, 2) for c in range(4)], 1)
meas = uv + rng.normal(0, 1.0, uv.shape)
w = np.where(rng.random(meas.shape[:-1]) < 0.9, 1 / 3.0, 0.0)
x0 = np.zeros_like(x_true); x0[:, :3] = x_true[:, :3] + 0.05
solver = build.SkelSolver(h, flat, meas, w, 1 / 120.0, loss="abs")
solver.solve(x0, max_iter=2)
torch.cuda.synchronize()
t0 = time.perf_counter()
x, info = solver.solve(x0, max_iter=200)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
err = np.linalg.norm(pose(x) - P3, axis=-1)
# evaluation rate on a large batch (device resident)
M = args.eval_frames
dev = torch.device("cuda", 0)
P = x.shape[1]
xb = torch.as_tensor(np.tile(x, (M // N + 1, 1))[:M]).to(dev)
mb = torch.as_tensor(np.tile(meas, (M // N + 1, 1, 1, 1))[:M]).to(dev)
wb = torch.as_tensor(np.tile(w, (M // N + 1, 1, 1))[:M]).to(dev)
cost = torch.zeros(M, dtype=torch.float64, device=dev); gg = torch.zeros(M, P, dtype=torch.float64, device=dev)
HH = torch.zeros(M, P * (P + 1) // 2, dtype=torch.float64, device=dev)
for _ in range(3):
    h.call_dev("acino_skel_eval_dev", M, xb, mb, wb, cost, gg, HH)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    h.call_dev("acino_skel_eval_dev", M, xb, mb, wb, cost, gg, HH)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# per-kernel time of one LM attempt (CUDA events, 5 repetitions, device resident)
def timed(name, *a):
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(5):
        h.call_dev(name, *a)
    a1.record(); torch.cuda.synchronize()
    return a0.elapsed_time(a1) / 5
st, tr = solver.st
phases = {}
phases["skel_eval"] = timed("acino_skel_eval_dev", N, st["x"], solver.meas, solver.w, st["cost"], st["g"], st["H"])
phases["skel_prepare"] = timed("acino_skel_prepare_dev", N, 1, st["x"], st["g"], solver.sw, solver.lo, solver.hi, st["gtot"], st["fixed"], st["cost_s"])
phases["skel_assemble"] = timed("acino_skel_assemble_dev", N, st["H"], st["gtot"], st["fixed"], solver.sw, 1e-3, solver.AB, solver.d)
def band():
    h.call_dev("acino_skel_assemble_dev", N, st["H"], st["gtot"], st["fixed"], solver.sw, 1e-3, solver.AB, solver.d)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    h.call_dev("acino_band_solve_dev", N * P, 3 * P, solver.AB, solver.d, solver.info)
    a1.record(); torch.cuda.synchronize()
    return a0.elapsed_time(a1)
phases["band_solve"] = float(np.median([band() for _ in range(5)]))
phases["skel_trial"] = timed("acino_skel_trial_dev", N, 1, st["x"], solver.d, solver.lo, solver.hi, tr["x"])
phases["skel_pred"] = timed("acino_skel_pred_dev", N, st["x"], tr["x"], st["gtot"], st["H"], solver.sw, solver.pred, solver.step)
bytes_per_frame = 8 * (P + 4 * 15 * 3 + 1 + P + P * (P + 1) // 2)
print(json.dumps({"frames": N, "P": P, "iters": info["iters"], "attempts": info["n_solve"], "seconds": dt,
                  "ms_per_attempt": 1e3 * dt / info["n_solve"], "F0": info["F0"], "F": info["F"], "converged": info["converged"],
                  "marker_median_err_m": float(np.median(err)), "eval_frames": M, "eval_ms": ms,
                  "eval_frames_per_s": M / ms * 1e3, "eval_GBps": M * bytes_per_frame / ms / 1e6,
                  "attempt_kernel_ms": {k: round(v, 4) for k, v in phases.items()}}))
