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


def lm_solve_rate(handle, n_frames):
    """LM iterations/sec of the full FTE solve (projected LM + block cyclic reduction) on one GPU."""
    import torch

    import synth
    from acinoset_b200 import lm

    def reproject(x):
        pos, uv = handle.fk_project(x.astype(np.float32))
        return pos.astype(np.float64), uv.astype(np.float64)

    p = synth.make_fte_problem(n_frames, None, None, seed=3, reproject=reproject)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"])
    sol.solve(p["x0"], max_iter=3)                      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, info = sol.solve(p["x0"], max_iter=60)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"frames": n_frames, "lm_iters_per_sec": info["n_solve"] / dt, "ms_per_iter": 1e3 * dt / info["n_solve"],
            "iters": info["n_solve"], "accepted": info["iters"], "seconds": dt, "converged": bool(info["converged"]),
            "what": "one iteration = fte_eval + assemble + block-cyclic-reduction solve + step acceptance (fp64 solve)"}


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
            "sample": f"{reps} pass(es) over {n} frames of the same batch, fp64 C restatement (oracle/c/fte_oracle.c), "
                      f"OpenMP over frames, {dt:.1f} s; reference Pyomo+IPOPT path not runnable in this image"}


def config_dict(n_gpus):
    return {"workload": f"fte_eval 6cam x 20kpt x {FRAMES_PER_SEQ}-frame sequences, {SEQS} sequences "
                        f"({SEQS * FRAMES_PER_SEQ} frames) per step per GPU, one launch/step",
            "cameras": C, "keypoints": L, "frames_per_sequence": FRAMES_PER_SEQ, "sequences_per_step_per_gpu": SEQS,
            "l2_policy": f"inputs+outputs {SEQS * FRAMES_PER_SEQ * BYTES_PER_FRAME / 1e6:.0f} MB per step > 126 MB L2",
            "parallelism": f"frames sharded over {n_gpus} GPU(s), no collective"}


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
                                   f"of the reference path (oracle/c/fte_oracle.c), OpenMP {cores} threads; the "
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

    # ---- single-sequence latency (1000 frames, back-to-back launches)
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
    single_us = 1e3 * e0.elapsed_time(e1) / reps

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

    # ---- secondary metric of BASELINE.json ("LM iters/sec"): full LM/FTE solve of configs[2] (10 000 frames),
    #      outside the timed region, rank 0 at N = 1 only
    lm_info = None
    if world == 1 and not args.no_lm:
        try:
            lm_info = lm_solve_rate(h, 10000)
        except Exception as e:      # a reported extra, never the measured path
            lm_info = {"error": str(e)}

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
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(world), single_sequence_1000f_us_per_launch=single_us,
                           campoint_pairs_per_sec=value * C * L, lm_solve_10000_frames=lm_info),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": f"of {peak_kind}",
                         "kernel": "fte_eval_kernel<true>", "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_frame": BYTES_PER_FRAME,
                         "note": "compute-bound above the fp32 ridge (see DESIGN.md); HBM fraction reported as mandated"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "acino_fte_eval (C ABI, pinned host buffers)"},
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
    if world > 1:
        dist.destroy_process_group()
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lm", action="store_true", help="skip the LM-solve secondary metric")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
