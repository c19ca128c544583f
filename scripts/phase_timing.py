"""Per-phase cycle counts of fte_eval (debug build with -DACINO_PHASE_TIMING at scratch/libacino_timing.so)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acinoset_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "..", "scratch", os.environ.get("ACINO_TIMING_LIB", "libacino_timing.so"))
L.lib = L._load()
import acinoset_b200 as ab, synth, torch
K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(0); h.set_cameras(K, D, R, t)
n = 256000
rng = np.random.default_rng(0)
x = synth.make_trajectory(n, rng).astype(np.float32)
pos, uv = h.fk_project(x)
meas = (uv + rng.normal(0, 2, uv.shape)).astype(np.float32); w = np.full((n, 6, 20), 0.2, np.float32)
dev = torch.device("cuda:0")
xd, md, wd = (torch.from_numpy(a).to(dev) for a in (x, meas, w))
c = torch.empty(n, device=dev); g = torch.empty(n, 25, device=dev); H = torch.empty(n, 325, device=dev)
for _ in range(3): h.fte_eval_dev(xd, md, wd, c, g, H)
torch.cuda.synchronize()
L.lib.acino_debug_phase_reset()
h.fte_eval_dev(xd, md, wd, c, g, H); torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
L.lib.acino_debug_phase_cycles(out)
cy = np.array(list(out)[:9], dtype=np.float64) / (n / 8)
names = ["-", "P0+P1a", "P1b", "P2 loop", "P2b", "P3", "P4", "next-tile issue", "P5"]
tot = cy.sum()
for nm, v in zip(names, cy): print(f"{nm:10s} {v:9.0f} cycles/CTA  {v / tot * 100:5.1f}%")
print("total", tot)
