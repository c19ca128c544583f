"""BASELINE.json configs[0]: 6 cam x 20 kpt x 90-frame pairwise DLT triangulation (reference: get_pairwise_3d_points_from_df,
0.52 s on CPU) - the drop-in DataFrame call and the dense kernel behind it, plus a 100 000-frame launch.
    python scripts/bench_tri.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth
from acinoset_b200 import calib, fte

g = np.load(os.path.join(ROOT, "tests", "golden", "triangulate.npz"))
f = np.load(os.path.join(ROOT, "tests", "golden", "fisheye.npz"))
K, D, R, t = f["K"], f["D"].reshape(-1, 4), f["R"], f["t"]
meas, lik = g["meas"], g["lik"]                                   # the 90-frame problem the reference was run on
df = synth.dense_to_long_df(meas, lik, fte.MARKERS)
dff = df[df["likelihood"] > 0.5].reset_index(drop=True)
import io, contextlib
def run_df():
    with contextlib.redirect_stdout(io.StringIO()):
        return calib.get_pairwise_3d_points_from_df(dff, K, D, R, t, calib.triangulate_points_fisheye)
run_df()
t0 = time.perf_counter(); out = run_df(); t_df = time.perf_counter() - t0
valid = lik > 0.5
calib.triangulate_pairwise_dense(meas, valid, K, D, R, t)
t0 = time.perf_counter(); pos, cnt = calib.triangulate_pairwise_dense(meas, valid, K, D, R, t); t_dense = time.perf_counter() - t0
ok = ~np.isnan(g["tri_pos"][..., 0])
err = float(np.abs(pos[ok] - g["tri_pos"][ok]).max())
reps = 100000 // 90 + 1
big = np.tile(meas, (reps, 1, 1, 1))[:100000]; bigv = np.tile(valid, (reps, 1, 1))[:100000]
calib.triangulate_pairwise_dense(big[:1000], bigv[:1000], K, D, R, t)
t0 = time.perf_counter(); calib.triangulate_pairwise_dense(big, bigv, K, D, R, t); t_big = time.perf_counter() - t0
print(json.dumps({"config": "6cam x 20kpt x 90 frames", "dataframe_call_s": t_df, "rows_out": int(len(out)),
                  "dense_call_s_host_buffers": t_dense, "max_abs_err_vs_reference_m": err,
                  "reference_cpu_s": 0.52, "frames_100000_host_buffers_s": t_big,
                  "frames_per_s_100000": 100000 / t_big}))
