#!/usr/bin/env python
"""bench.py - residual+Jacobian evaluations/sec of the FTE hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 6 cameras x 20 keypoints x 1000-frame sequences, fp32.
One *step* = one pass of `fte_eval` over a batch of SEQS independent 1000-frame sequences
(one kernel launch over SEQS*1000 frames).  A single 1000-frame sequence moves 2.9 MB and
lives entirely in L2 / below launch latency, so the batch is sized so that the step's
inputs+outputs (SEQS x 2.9 MB) exceed the 126 MB L2 several times over (timing rule:
"inputs larger than L2"); the single-sequence launch latency is reported next to it.
Weak scaling: every rank evaluates its own SEQS sequences, no data-path collective (frames
are independent for evaluation).

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_SEQ = 1000
SEQS = 256
C, L, NA, NU = 6, 20, 25, 325
BYTES_PER_FRAME = 4 * (3 * C * L + 2 * NA + NU + 1)   # 2944 B algorithmic
# fp32 operations per frame counted from the kernel source (csrc/fte_eval.cu): projection + Jacobian +
# two loss evaluations = 275 flops per (camera, marker) x 120, + FK 0.5 k + spatial inertias 1.2 k + subtree sums 0.5 k +
# y = I tau 1.8 k + block dot products 2.0 k
FLOPS_PER_FRAME = 39000
FP32_PEAK_TFLOPS = 72.4
METRIC = "fte_residual_jacobian_evals_per_sec"
UNIT = "frames/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def make_inputs_gpu(handle, n_frames, seed):
    """Synthetic batch generated with the product's own fk_project kernel."""
    import synth

    def reproject(x):
        pos, uv = handle.fk_project(x.astype(np.float32))
        return pos.astype(np.float64), uv.astype(np.float64)

    p = synth.make_fte_problem(n_frames, None, None, seed=seed, reproject=reproject)
    return (p["x0"].astype(np.float32), p["meas"].astype(np.float32), p["w"].astype(np.float32))


def make_inputs_cpu(n_frames, seed):
    """Same generator with the oracle's FK / projection (reference arm: no GPU involved)."""
    import synth
    from oracle import fisheye, skeleton

    p = synth.make_fte_problem(n_frames, skeleton.cheetah_fk_active, fisheye.project, seed=seed)
    return p["x0"].astype(np.float32), p["meas"].astype(np.float32), p["w"].astype(np.float32), p["cams"]


def _reprojector(handle):
    def reproject(x):
        pos, uv = handle.fk_project(x.astype(np.float32))
        return pos.astype(np.float64), uv.astype(np.float64)
    return reproject


def _max_over_ranks(v, dev, world):
    import torch
    import torch.distributed as dist

    tt = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


def lm_solve_block(handle, n_frames, rank, world, dev, with_rms=True, max_iter=60):
    """LM iterations/sec of the FTE solve (replaces opt.solve, all_optimizations.py:503-524) with the frames
    sharded over the ranks (BASELINE.json configs[2] at world = 1 / 10 000 frames, configs[4] at 100 000 frames).
    Every rank generates only its own shard of the synthetic trajectory.  Time = max over ranks."""
    import torch
    import torch.distributed as dist

    import synth
    from acinoset_b200 import lm

    f0, n = lm.shard_frames(n_frames, world)[rank]
    p = synth.make_fte_problem(n, None, None, seed=3, reproject=_reprojector(handle), start=f0)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"], frame0=f0, n_global=n_frames, rank=rank, world=world)
    sol.solve(p["x0"], max_iter=3)                      # warm-up (lazy module load, NCCL channels, graph capture)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    x, info = sol.solve(p["x0"], max_iter=max_iter)
    torch.cuda.synchronize()
    dt = _max_over_ranks(time.perf_counter() - t0, dev, world)
    out = {"frames": n_frames, "gpus": world, "frames_per_gpu": n, "lm_iters_per_sec": info["n_solve"] / dt,
           "attempts_per_sec": info["n_solve"] / dt, "ms_per_attempt": 1e3 * dt / info["n_solve"],
           "ms_per_attempt_steady": _max_over_ranks(info.get("attempt_ms_steady") or 0.0, dev, world),
           "ms_per_attempt_how": "ms_per_attempt = wall time of the whole solve (first eager attempt and graph capture "
                                 "included) / attempts; _steady = median device time (CUDA events) of the graph-replayed "
                                 "attempts, max over ranks",
           "attempts": info["n_solve"], "accepted_iterations": info["iters"], "seconds_to_converge": dt,
           "F": info["F"], "converged": bool(info["converged"]), "bcr_info": info["bcr_info"],
           "collectives_per_attempt": info.get("collectives_per_attempt"), "cuda_graph": info.get("graph"),
           "collective_us_per_attempt": sol.exchange_us() if world > 1 else 0.0,
           "collective_us_how": "the attempt's two all_gathers (22 800 + 8 doubles per rank) timed back to back with CUDA "
                                "events outside the solve, 50 repetitions",
           "host_syncs_per_attempt": info.get("host_syncs_per_attempt", 1),
           "what": "one attempt = assemble + block-cyclic-reduction solve (fp64) + interface exchange + trial fte_eval "
                   "+ acceptance test"}
    if with_rms:
        pos, _ = handle.fk_project(x.astype(np.float32))
        post, _ = handle.fk_project(p["x_true"].astype(np.float32))
        se = torch.tensor([float(((pos - post) ** 2).sum()), float(pos.shape[0] * pos.shape[1])], device=dev,
                          dtype=torch.float64)
        if world > 1:
            dist.all_reduce(se)
        out["marker_rms_m"] = float(np.sqrt(se[0].item() / se[1].item()))
    sol.close()
    del sol
    torch.cuda.empty_cache()
    return out


def lm_parity_block(handle, rank, world, dev, n_frames=600):
    """Sharded vs single-rank solve of the SAME problem, inside the run (the driver's pytest box has one GPU)."""
    import torch
    import torch.distributed as dist

    import synth
    from acinoset_b200 import lm

    p = synth.make_fte_problem(n_frames, None, None, seed=17, reproject=_reprojector(handle))
    shards = lm.shard_frames(n_frames, world)
    f0, n = shards[rank]
    sol = lm.FTESolver(handle, p["meas"][f0:f0 + n], p["w"][f0:f0 + n], p["Ts"], frame0=f0, n_global=n_frames,
                       rank=rank, world=world)
    x, info = sol.solve(p["x0"][f0:f0 + n], max_iter=40)
    xs = [torch.zeros(shards[r][1], NA, dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(xs, torch.from_numpy(np.ascontiguousarray(x)).to(dev))
    res = torch.zeros(3, dtype=torch.float64, device=dev)
    if rank == 0:
        xg = torch.cat(xs).cpu().numpy()
        x1, info1 = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"]).solve(p["x0"], max_iter=40)
        res[0] = abs(info["F"] - info1["F"]) / abs(info1["F"])
        res[1] = float(np.abs(xg - x1).max())
        res[2] = info1["iters"]
    dist.broadcast(res, 0)
    sol.close()
    dF, dx = float(res[0].item()), float(res[1].item())
    return {"frames": n_frames, "dF_rel": dF, "max_dx": dx, "iters_sharded": info["iters"], "iters_single": int(res[2].item()),
            "ok": bool(dF < 1e-6 and dx < 1e-3 and info["bcr_info"] == 0), "tolerance": "dF_rel < 1e-6, max|dx| < 1e-3"}


def _sba_problem(handle, views, seed):
    import synth

    return synth.make_sba_problem(views, lambda X, k, d, r, tt: handle.project_points(X, k, d, r, tt), seed=seed)


def _sba_start(p):
    """Initial 3-D points: truth + 2 cm noise (what pairwise triangulation with the perturbed extrinsics gives)."""
    n_pts = len(p["points_3d_true"])
    return (p["points_3d_true"] + np.random.default_rng(3).normal(0, 0.02, (n_pts, 3))).astype(np.float64)


def sba_block(handle, rank, world, dev, views=5000):
    """BASELINE.json configs[3]: 6 cameras x `views` checkerboard views, extrinsics + points, views sharded over the
    ranks (one all_reduce of the reduced 36 x 36 camera system per LM attempt; calib.py:362-390)."""
    import torch
    import torch.distributed as dist

    from acinoset_b200 import sba

    p = _sba_problem(handle, views, seed=4)
    n_pts = len(p["points_3d_true"])
    pidx, cidx = p["point_3d_indices"], p["camera_indices"]
    pts0 = _sba_start(p)
    x0 = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]]), p["t0"].ravel()])
    p0, npl, ids = sba.shard_points(pidx, n_pts, world)[rank]
    prob = sba.SBAProblem(p["points_2d"][ids], pidx[ids] - p0, cidx[ids], p["K"], p["D"], npl, device=handle.device,
                          rank=rank, world=world)
    s = prob.st[0]
    s["params"].copy_(torch.as_tensor(x0).to(dev))
    s["pts"].copy_(torch.as_tensor(pts0[p0:p0 + npl]).to(dev))
    for _ in range(3):
        prob._eval(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prob._eval(s)
    e1.record()
    torch.cuda.synchronize()
    ms_eval = _max_over_ranks(e0.elapsed_time(e1) / 20, dev, world)
    prob.solve(x0, pts0[p0:p0 + npl], max_nfev=3, ftol=1e-10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = prob.solve(x0, pts0[p0:p0 + npl], max_nfev=200, ftol=1e-10)
    torch.cuda.synchronize()
    dt = _max_over_ranks(time.perf_counter() - t0, dev, world)
    n_obs = len(pidx)
    blk = {"views": int(p["n_views"]), "n_obs": int(n_obs), "n_pts": int(n_pts), "gpus": world,
           "obs_per_sec": n_obs / (ms_eval * 1e-3), "eval_ms": ms_eval,
           "eval_GBps_at_176B_per_obs": n_obs * 176 / (ms_eval * 1e-3) / 1e9, "lm_iters_per_sec": out["nfev"] / dt,
           "solve_s": dt, "nfev": out["nfev"], "initial_cost": out["cost0"], "final_cost": out["cost"],
           "status": out["status"]}
    # gauge-free extrinsic error vs the ground truth (relative pose camera 0 -> camera c)
    _, r_new, t_new = sba.params_to_points_extrinsics(np.concatenate([out["params"], np.zeros(3)]), 6, 1)

    def rel(Rs, ts):
        return [(Rs[c] @ Rs[0].T, np.reshape(ts[c], 3) - Rs[c] @ Rs[0].T @ np.reshape(ts[0], 3)) for c in range(6)]

    def ang(A, B):
        return float(np.degrees(np.arccos(np.clip((np.trace(A @ B.T) - 1) / 2, -1, 1))))

    rt, ro = rel(p["R_true"], p["t_true"]), rel(r_new, t_new)
    blk["max_rel_rot_err_deg"] = max(ang(a[0], b[0]) for a, b in zip(rt, ro))
    del prob
    torch.cuda.empty_cache()
    return blk, p, x0, pts0


def sba_parity_block(handle, rank, world, dev, views=300):
    """View-sharded vs single-rank SBA solve of the same problem (tests/_mgpu_worker_sba.py, inside the run)."""
    import torch
    import torch.distributed as dist

    from acinoset_b200 import sba

    p = _sba_problem(handle, views, seed=21)
    n_pts = len(p["points_3d_true"])
    pidx, cidx = p["point_3d_indices"], p["camera_indices"]
    pts0 = _sba_start(p)
    x0 = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]]), p["t0"].ravel()])
    p0, npl, ids = sba.shard_points(pidx, n_pts, world)[rank]
    prob = sba.SBAProblem(p["points_2d"][ids], pidx[ids] - p0, cidx[ids], p["K"], p["D"], npl, device=handle.device,
                          rank=rank, world=world)
    out = prob.solve(x0, pts0[p0:p0 + npl], max_nfev=1000, ftol=1e-10)
    res = torch.zeros(2, dtype=torch.float64, device=dev)
    if rank == 0:
        one = sba.SBAProblem(p["points_2d"], pidx, cidx, p["K"], p["D"], n_pts, device=handle.device).solve(
            x0, pts0, max_nfev=1000, ftol=1e-10)
        res[0] = abs(out["cost"] - one["cost"]) / abs(one["cost"])
        res[1] = float(np.abs(out["params"] - one["params"]).max())
    dist.broadcast(res, 0)
    dF, dC = float(res[0].item()), float(res[1].item())
    return {"views": views, "dF_rel": dF, "max_d_cam": dC, "ok": bool(dF < 1e-8 and dC < 1e-6),
            "tolerance": "dF_rel < 1e-8, max|d camera params| < 1e-6"}


def tri_block(handle):
    """BASELINE.json configs[0]: 6 cam x 20 kpt x 90-frame pairwise DLT triangulation (calib.py:394-423) on the
    committed golden problem; parity against the reference's own output (tests/golden/triangulate.npz)."""
    from acinoset_b200 import calib

    g = np.load(os.path.join(ROOT, "tests", "golden", "triangulate.npz"))
    f = np.load(os.path.join(ROOT, "tests", "golden", "fisheye.npz"))
    K, D, R, t = f["K"], f["D"].reshape(-1, 4), f["R"], f["t"]
    meas, valid = g["meas"], g["lik"] > 0.5
    calib.triangulate_pairwise_dense(meas, valid, K, D, R, t, device=handle.device)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        pos, cnt = calib.triangulate_pairwise_dense(meas, valid, K, D, R, t, device=handle.device)
    gpu_s = (time.perf_counter() - t0) / reps
    ok = ~np.isnan(g["tri_pos"][..., 0])
    err = float(np.abs(pos[ok] - g["tri_pos"][ok]).max())
    nan_equal = bool(np.array_equal(np.isnan(pos[..., 0]), np.isnan(g["tri_pos"][..., 0])))
    from oracle import triangulate as o_tri

    t0 = time.perf_counter()
    o_tri.pairwise_mean_dense(meas, valid, K, D, R, t)
    cpu_s = time.perf_counter() - t0
    return {"frames": int(meas.shape[0]), "gpu_s": gpu_s, "api": "calib.triangulate_pairwise_dense (host buffers in/out)",
            "max_abs_err_m_vs_reference_golden": err, "nan_pattern_equal": nan_equal,
            "cpu_baseline": {"value": cpu_s, "unit": "s", "cores": 1, "kind": "port",
                             "sample": "oracle.triangulate.pairwise_mean_dense (NumPy fp64 restatement of calib.py:394-423) on "
                                       "the same 90 frames; the reference itself measured 0.52 s in the build container "
                                       "(BASELINE.md section 2) and cannot run on the GPU box"}}


def sba_cpu_baseline(p, x0, pts0, target_obs=200000):
    """Residual evaluation of the NumPy restatement (oracle.sba.cost_func_points_extrinsics) on a sub-sample."""
    from oracle import sba as o_sba

    pidx, cidx = p["point_3d_indices"], p["camera_indices"]
    keep = pidx < (target_obs // 2 // 54 + 1) * 54 if len(pidx) > target_obs else np.ones(len(pidx), bool)
    n_pts = int(pidx[keep].max()) + 1
    params = np.concatenate([x0, pts0[:n_pts].ravel()])
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        o_sba.cost_func_points_extrinsics(params, 6, n_pts, pidx[keep], cidx[keep], p["K"], p["D"],
                                          p["points_2d"][keep].astype(np.float64))
    dt = (time.perf_counter() - t0) / reps
    n = int(keep.sum())
    return {"value": n / dt, "unit": "obs/s", "cores": 1, "kind": "port",
            "sample": f"{n} observations, residuals only, vectorised NumPy fp64 restatement of calib.py:355-359 "
                      "(the reference calls cv2.fisheye.projectPoints once per observation: 3.3e4 obs/s, BASELINE.md)"}


def lm_cpu_baseline(handle, n_frames=100, target=10000):
    """The same LM algorithm in fp64 NumPy + SciPy sparse LU (oracle/lm.py) on a small problem, extrapolated."""
    import synth
    from oracle import lm as o_lm

    p = synth.make_fte_problem(n_frames, None, None, seed=3, reproject=_reprojector(handle))
    t0 = time.perf_counter()
    _, info = o_lm.solve(p, p["x0"], max_iter=4)
    dt = time.perf_counter() - t0
    per_attempt = dt / max(info["n_eval"] - 1, 1)
    return {"value": 1.0 / (per_attempt * target / n_frames), "unit": "LM iterations/s", "cores": 1, "kind": "port",
            "sample": f"{info['n_eval'] - 1} attempts on {n_frames} frames ({dt:.1f} s), EXTRAPOLATED linearly to {target} "
                      "frames; the reference's Pyomo + IPOPT solve cannot run in this image"}


def copy_ceiling(bufs_in, bufs_out, dev, reps=3):
    """Concurrent H2D + D2H of the e2e call's own buffers on two streams: the PCIe bound of the host-buffer API."""
    import torch

    din = [torch.empty(b.shape, dtype=torch.float32, device=dev) for b in bufs_in]
    dout = [torch.empty(b.shape, dtype=torch.float32, device=dev) for b in bufs_out]
    tin = [torch.from_numpy(b) for b in bufs_in]
    tout = [torch.from_numpy(b) for b in bufs_out]
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            for d, h_ in zip(din, tin):
                d.copy_(h_, non_blocking=True)
        with torch.cuda.stream(s2):
            for d, h_ in zip(dout, tout):
                h_.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def host_threads():
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1: ask the OS instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_time(x, meas, w, cams, target_s=12.0, threads=0):
    """Time the oracle C port (oracle/c/fte_oracle.c) on a bounded sample; returns dict."""
    from oracle import c_port

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    K, D, R, t, _ = cams
    threads = host_threads() if threads <= 0 else threads
    cores = threads
    n0 = min(2000, x.shape[0])
    t0 = time.perf_counter()
    c_port.fte_eval(x[:n0], meas[:n0], w[:n0], K, D, R, t, n_threads=threads)
    rate0 = n0 / (time.perf_counter() - t0)
    n = int(min(x.shape[0], max(n0, rate0 * target_s)))
    reps = max(1, int(round(rate0 * target_s / n)))           # the whole batch several times over when the host is fast
    t0 = time.perf_counter()
    for _ in range(reps):
        c_port.fte_eval(x[:n], meas[:n], w[:n], K, D, R, t, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} pass(es) over {n} frames of the same batch, fp64 C restatement (oracle/c/fte_oracle.c, gcc -O2 without -march=native), "
                      f"OpenMP over frames, {dt:.1f} s; reference Pyomo+IPOPT path not runnable in this image"}


def config_dict(n_gpus):
    return {"workload": f"fte_eval 6cam x 20kpt x {FRAMES_PER_SEQ}-frame sequences, {SEQS} sequences "
                        f"({SEQS * FRAMES_PER_SEQ} frames) per step per GPU, one launch/step",
            "cameras": C, "keypoints": L, "frames_per_sequence": FRAMES_PER_SEQ, "sequences_per_step_per_gpu": SEQS,
            "l2_policy": f"inputs+outputs {SEQS * FRAMES_PER_SEQ * BYTES_PER_FRAME / 1e6:.0f} MB per step > 126 MB L2",
            "parallelism": f"fte_eval: frames sharded over {n_gpus} GPU(s), no collective (frames are independent); the "
                           f"sharded LM solve (config.lm_sharded_100000_frames) exchanges interface blocks every attempt"}


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU restatement of the reference path, all host threads."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import c_port

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    cores = host_threads()
    sample_seqs = 16
    n = sample_seqs * FRAMES_PER_SEQ
    x, meas, w, cams = make_inputs_cpu(n, seed=1000)
    K, D, R, t, _ = cams
    xd, md, wd = x.astype(np.float64), meas.astype(np.float64), w.astype(np.float64)
    for _ in range(args.warmup):
        c_port.fte_eval(xd, md, wd, K, D, R, t, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_port.fte_eval(xd, md, wd, K, D, R, t, n_threads=cores)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {sample_seqs} of the {SEQS} sequences ({n} frames), fp64 C restatement "
                                   f"of the reference path (oracle/c/fte_oracle.c, gcc -O2 without -march=native), OpenMP {cores} threads; the "
                                   f"reference's own Pyomo+IPOPT evaluation cannot run here (pyomo/ipopt absent)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import acinoset_b200 as ab
    import synth

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, D, R, t, _ = synth.load_dummy_scene()
    h = ab.Handle(local)
    h.set_cameras(K, D, R, t)

    n = SEQS * FRAMES_PER_SEQ
    x, meas, w = make_inputs_gpu(h, n, seed=1000 + rank)
    xd, md, wd = (torch.from_numpy(a).to(dev) for a in (x, meas, w))
    cost = torch.empty(n, device=dev)
    g = torch.empty(n, NA, device=dev)
    H = torch.empty(n, NU, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, one launch each, CUDA events on the launch stream
    for _ in range(max(args.warmup, 3)):
        h.fte_eval_dev(xd, md, wd, cost, g, H)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = h.launch_count
    barrier()
    evs[0].record()
    for i in range(args.steps):
        h.fte_eval_dev(xd, md, wd, cost, g, H)
        evs[i + 1].record()
    barrier()
    launches = h.launch_count - l0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # ---- single-sequence latency (configs[1] literally: ONE 1000-frame sequence per launch).  Device time per launch from
    #      a CUDA graph of 50 back-to-back launches (what the LM loop does); the eager Python call rate is reported beside it
    xs, ms_, ws = xd[:FRAMES_PER_SEQ], md[:FRAMES_PER_SEQ], wd[:FRAMES_PER_SEQ]
    cs, gs, Hs = cost[:FRAMES_PER_SEQ], g[:FRAMES_PER_SEQ], H[:FRAMES_PER_SEQ]
    for _ in range(20):
        h.fte_eval_dev(xs, ms_, ws, cs, gs, Hs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        h.fte_eval_dev(xs, ms_, ws, cs, gs, Hs)
    e1.record()
    torch.cuda.synchronize()
    single_eager_us = 1e3 * e0.elapsed_time(e1) / reps
    single_us = single_eager_us
    try:
        per_graph = 50
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(per_graph):
                h.fte_eval_dev(xs, ms_, ws, cs, gs, Hs)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        single_us = 1e3 * e0.elapsed_time(e1) / (10 * per_graph)
        del gr
    except Exception:
        pass

    # ---- end to end through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H per step
    hx, hm, hw = (torch.from_numpy(a).pin_memory().numpy() for a in (x, meas, w))
    hc = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    hg = torch.empty(n, NA, dtype=torch.float32).pin_memory().numpy()
    hH = torch.empty(n, NU, dtype=torch.float32).pin_memory().numpy()
    e2e_steps = max(2, min(args.steps, 5))
    h.fte_eval(hx, hm, hw, out=(hc, hg, hH))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.fte_eval(hx, hm, hw, out=(hc, hg, hH))
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * n * e2e_steps / float(te.item())
    h2d = int(hx.nbytes + hm.nbytes + hw.nbytes)
    d2h = int(hc.nbytes + hg.nbytes + hH.nbytes)

    # ---- PCIe ceiling of the host-buffer call: the same buffers, H2D and D2H concurrently, no kernel
    ceil_dt = copy_ceiling((hx, hm, hw), (hc, hg, hH), dev)
    ceil_dt = _max_over_ranks(ceil_dt, dev, world)
    pcie_bound = world * n / ceil_dt
    del hx, hm, hw, hc, hg, hH

    # ---- everything below is OUTSIDE the timed region: the solver-level half of BASELINE.json's metric
    #      ("LM iters/sec") and the other configs, each with its own parity check and CPU number
    extras = {}
    del xd, md, wd, cost, g, H, xs, ms_, ws, cs, gs, Hs
    torch.cuda.empty_cache()

    def guarded(name, fn):
        try:
            extras[name] = fn()
        except Exception as e:      # reported extras, never the measured path
            import traceback

            extras[name] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}

    if not args.no_lm:
        if world == 1:
            guarded("lm_solve_10000_frames", lambda: lm_solve_block(h, 10000, rank, world, dev))
        # the north-star multi-GPU path (configs[4]): 100 000 frames sharded over the ranks, at EVERY N
        guarded("lm_sharded_100000_frames", lambda: lm_solve_block(h, args.lm_frames, rank, world, dev))
        if world > 1:
            guarded("lm_sharded_parity_vs_single_rank", lambda: lm_parity_block(h, rank, world, dev))
            if isinstance(extras.get("lm_sharded_100000_frames"), dict):
                extras["lm_sharded_100000_frames"]["parity_vs_single_rank"] = extras.get("lm_sharded_parity_vs_single_rank")
    sba_prob = None
    if not args.no_sba:
        def _sba():
            nonlocal sba_prob
            blk, p_, x0_, pts0_ = sba_block(h, rank, world, dev, args.sba_views)
            sba_prob = (p_, x0_, pts0_)
            return blk
        guarded("sba_config3", _sba)
        if world > 1:
            guarded("sba_sharded_parity_vs_single_rank", lambda: sba_parity_block(h, rank, world, dev))
            if isinstance(extras.get("sba_config3"), dict):
                extras["sba_config3"]["parity_vs_single_rank"] = extras.get("sba_sharded_parity_vs_single_rank")
    if world == 1 and rank == 0:
        guarded("tri_config0", lambda: tri_block(h))
        if not args.no_cpu_baseline:
            if sba_prob is not None and isinstance(extras.get("sba_config3"), dict):
                try:
                    extras["sba_config3"]["cpu_baseline"] = sba_cpu_baseline(*sba_prob)
                except Exception as e:
                    extras["sba_config3"]["cpu_baseline"] = {"error": str(e)}
            if isinstance(extras.get("lm_solve_10000_frames"), dict) and "error" not in extras["lm_solve_10000_frames"]:
                try:
                    extras["lm_solve_10000_frames"]["cpu_baseline"] = lm_cpu_baseline(h)
                except Exception as e:
                    extras["lm_solve_10000_frames"]["cpu_baseline"] = {"error": str(e)}

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        kern_ms = float(np.mean(per_launch_ms))
        achieved = n * BYTES_PER_FRAME / (kern_ms * 1e-3) / 1e9
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "fte_eval_traffic.json")
        if os.path.exists(tfile):
            try:
                traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        tflops = n * FLOPS_PER_FRAME / (kern_ms * 1e-3) / 1e12
        cfg = dict(config_dict(world), single_sequence_1000f_us_per_launch=single_us,
                   single_sequence_1000f_us_per_eager_python_call=single_eager_us,
                   campoint_pairs_per_sec=value * C * L)
        cfg.update(extras)
        if world > 1:
            cfg["parallelism"] += ("; e2e: all ranks share one host (pinned buffers on one NUMA node): host-memory/PCIe "
                                   "bound, see e2e.pcie_bound_frames_per_sec")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": f"of {peak_kind}",
                         "kernel": h.fte_kernel_name(), "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_frame": BYTES_PER_FRAME,
                         "fp32": {"flops_per_frame": FLOPS_PER_FRAME, "achieved_tflops": tflops, "peak_tflops": FP32_PEAK_TFLOPS,
                                  "frac": tflops / FP32_PEAK_TFLOPS,
                                  "peak_source": "measured FFMA/FFMA2 peak on this pool's B200 (scripts/micro/ffma2_bench.cu)"},
                         "note": "issue/latency-bound on the CUDA cores above the fp32 ridge (DESIGN.md 4.1); HBM fraction "
                                 "reported as mandated, fp32 fraction beside it"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "acino_fte_eval (C ABI, pinned host buffers)",
                    "pcie_bound_frames_per_sec": pcie_bound, "frac_of_pcie_bound": e2e_val / pcie_bound,
                    "pcie_bound_how": "the call's own pinned buffers copied H2D and D2H concurrently on two streams, no "
                                      "kernel, best of 3, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cams = synth.load_dummy_scene()
                out["cpu_baseline"] = cpu_baseline_time(x, meas, w, cams)
            except Exception as e:  # the baseline is a reported extra, never the measured path
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(out), flush=True)
    h.close()
    if world > 1:
        # every rank is done (the JSON line is out): leave without NCCL's communicator teardown, which has been seen to block
        # after graph-captured collectives (tests/_mgpu_worker.py) - a hung bench would cost the whole scaling record
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lm", action="store_true", help="skip the LM-solve blocks (configs[2], configs[4])")
    ap.add_argument("--no-sba", action="store_true", help="skip the SBA block (configs[3])")
    ap.add_argument("--lm-frames", type=int, default=100000, help="frames of the sharded LM solve (configs[4])")
    ap.add_argument("--sba-views", type=int, default=5000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
