"""BASELINE.json configs[4]: FTE LM solve with the frames sharded over the GPUs of one box (torchrun), one all_gather of
the interface super-blocks per LM iteration.  Prints iterations/s (max over ranks) on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_lm_sharded.py --frames 100000
"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100000)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import acinoset_b200 as ab
import synth
from acinoset_b200 import lm

K, D, R, t, _ = synth.load_dummy_scene()
h = ab.Handle(local); h.set_cameras(K, D, R, t)
N = args.frames
f0, n = lm.shard_frames(N, world)[rank]
def reproject(x):
    pos, uv = h.fk_project(x.astype(np.float32)); return pos.astype(np.float64), uv.astype(np.float64)
# every rank generates only its shard (same trajectory generator, shard start offset)
p = synth.make_fte_problem(n, None, None, seed=3, reproject=reproject, start=f0)
sol = lm.FTESolver(h, p["meas"], p["w"], p["Ts"], frame0=f0, n_global=N, rank=rank, world=world)
sol.solve(p["x0"], max_iter=3)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
x, info = sol.solve(p["x0"], max_iter=60)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}", dtype=torch.float64)
if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
pos, _ = h.fk_project(x.astype(np.float32)); post, _ = h.fk_project(p["x_true"].astype(np.float32))
se = torch.tensor([float(((pos - post) ** 2).sum()), float(pos.shape[0] * pos.shape[1])], device=f"cuda:{local}", dtype=torch.float64)
if world > 1: dist.all_reduce(se)
if rank == 0:
    print(json.dumps({"frames": N, "gpus": world, "frames_per_gpu": n, "iters": info["iters"], "attempts": info["n_solve"],
                      "seconds": float(dt.item()), "lm_iters_per_sec": info["n_solve"] / float(dt.item()),
                      "ms_per_attempt": 1e3 * float(dt.item()) / info["n_solve"], "F": info["F"], "converged": info["converged"],
                      "marker_rms_m": float(np.sqrt(se[0].item() / se[1].item())), "bcr_info": info["bcr_info"]}))
sol.close()
if world > 1: dist.destroy_process_group()
