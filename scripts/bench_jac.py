"""fte_jac (dense measurement Jacobian, SURVEY 8f-1 / the dense-J-out variant of 8d): frames/s and achieved HBM GB/s.
Algorithmic bytes per frame: read state 100 B; write pixels C*L*2*4 = 960 B and J C*L*2*25*4 = 24 000 B => 25 060 B."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import acinoset_b200 as ab
import torch

K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(0); h.set_cameras(K, D, R, t)
n = 64000                                     # 1.6 GB of Jacobian per launch (> 126 MB L2)
rng = np.random.default_rng(0)
x = torch.from_numpy(synth.make_trajectory(n, rng).astype(np.float32)).cuda()
uv = torch.empty(n, 6, 20, 2, device="cuda"); J = torch.empty(n, 6, 20, 2, 25, device="cuda")
for _ in range(3): h.fte_jac_dev(x, uv, J)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps): h.fte_jac_dev(x, uv, J)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
B = 4 * (25 + 6 * 20 * 2 + 6 * 20 * 2 * 25)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
print(json.dumps({"kernel": "fte_jac_kernel", "frames": n, "ms_per_launch": ms, "frames_per_sec": n / (ms * 1e-3),
                  "bytes_per_frame": B, "achieved_GBps": n * B / (ms * 1e-3) / 1e9, "peak_GBps": peak,
                  "frac": n * B / (ms * 1e-3) / 1e9 / peak}))
