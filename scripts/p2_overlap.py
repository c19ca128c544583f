"""Are the co-resident CTAs of fte_eval inside the camera loop (P2) at the same time?

Debug build with -DACINO_PHASE_TIMING at scratch/libacino_timing.so: thread 0 of every CTA records clock64 at the start and
the end of the camera loop of its first 64 tiles, and its %smid.  clock64 is per SM, so the intervals of the CTAs that share
an SM are comparable.  Prints, over all SMs, the share of the time during which k = 0..4 CTAs are inside the loop, and the
mean loop / rest duration; a lockstep wave shows up as k in {0, 4} only, a staggered one as k ~ 2.
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acinoset_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "..", "scratch", sys.argv[1] if len(sys.argv) > 1 else "libacino_timing.so")
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
L.lib.acino_debug_p2_trace_reset()
h.fte_eval_dev(xd, md, wd, c, g, H); torch.cuda.synchronize()
trace = np.zeros(1024 * 128 * 2, np.int64); smid = np.zeros(1024, np.int32)
L.lib.acino_debug_p2_trace(trace.ctypes.data_as(ctypes.c_void_p), smid.ctypes.data_as(ctypes.c_void_p))
trace = trace.reshape(1024, 128, 2)
n_cta = int(sys.argv[2]) if len(sys.argv) > 2 else 592
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/p2_trace_%s.npz" % os.path.basename(L.LIB_PATH), trace=trace[:n_cta], smid=smid[:n_cta])
n_done = (trace[:n_cta, :, 1] != 0).sum(axis=1)          # tiles every CTA processed
last_end = np.array([trace[b, n_done[b] - 1, 1] - trace[b, 0, 0] for b in range(n_cta)])
print("tiles per CTA: min %d mean %.1f max %d;  span first loop start -> last loop end (kcycles): min %.0f mean %.0f max %.0f" % (
    n_done.min(), n_done.mean(), n_done.max(), last_end.min() / 1e3, last_end.mean() / 1e3, last_end.max() / 1e3))
per_sm_last = []; per_sm_first = []
for sm in np.unique(smid[:n_cta]):
    ctas = np.nonzero(smid[:n_cta] == sm)[0]
    ends = [trace[b, n_done[b] - 1, 1] for b in ctas]; t0 = min(trace[b, 0, 0] for b in ctas)
    per_sm_last.append((max(ends) - t0) / 1e3); per_sm_first.append((min(ends) - t0) / 1e3)
print("per SM: first CTA done at %.0f kcycles (mean), last CTA done at %.0f (mean), %.0f (max)" % (np.mean(per_sm_first), np.mean(per_sm_last), np.max(per_sm_last)))
tiles = int(min(128, n_done.min()))
hist = np.zeros(8); loop_d = []; rest_d = []
per_tile_hist = np.zeros((tiles, 8))
for sm in np.unique(smid[:n_cta]):
    ctas = np.nonzero(smid[:n_cta] == sm)[0]
    ev = []
    for b in ctas:
        for it in range(tiles):
            s, e = trace[b, it]
            ev.append((s, +1, it)); ev.append((e, -1, it))
            loop_d.append(e - s)
            if it + 1 < tiles: rest_d.append(trace[b, it + 1, 0] - e)
    ev.sort()
    k = 0
    for (t0, d, it), (t1, _, _) in zip(ev[:-1], ev[1:]):
        k += d
        hist[k] += t1 - t0
        per_tile_hist[it, k] += t1 - t0
print("CTAs per SM:", np.bincount(np.bincount(smid[:n_cta])))
print("share of the time with k CTAs of the SM inside the camera loop (k = 0..4):", np.round(hist[:5] / hist.sum(), 3))
print("camera loop: mean %.0f cycles (p10 %.0f, p90 %.0f); rest of a tile: mean %.0f (p10 %.0f, p90 %.0f)" % (
    np.mean(loop_d), np.percentile(loop_d, 10), np.percentile(loop_d, 90), np.mean(rest_d), np.percentile(rest_d, 10), np.percentile(rest_d, 90)))
for it in (0, 1, 2, 5, 10, 20, 40, tiles - 1):
    if it < tiles:
        r = per_tile_hist[it]
        print("  events of tile %2d: k share" % it, np.round(r[:5] / max(r.sum(), 1), 2))
# one SM in detail
sm = smid[0]
ctas = np.nonzero(smid[:n_cta] == sm)[0]
t00 = trace[ctas, 0, 0].min()
for b in ctas:
    print("  SM %d CTA %4d loop start/end (kcycles):" % (sm, b), " ".join("%.1f-%.1f" % ((trace[b, it, 0] - t00) / 1e3, (trace[b, it, 1] - t00) / 1e3) for it in range(min(tiles, 8))))
