"""torchrun worker: SBA with the board views sharded over WORLD_SIZE GPUs (one all_reduce of the
reduced camera system per LM attempt) vs the single-GPU solve (rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import synth
    from acinoset_b200 import sba
    from oracle import fisheye

    n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    p = synth.make_sba_problem(n_views, fisheye.project, seed=21)
    K, D = p["K"], p["D"]
    pidx, cidx = p["point_3d_indices"], p["camera_indices"]
    n_pts = len(p["points_3d_true"])
    pts0 = (p["points_3d_true"] + np.random.default_rng(3).normal(0, 0.02, (n_pts, 3))).astype(np.float32)
    # the reference-named entry point: sharded because a process group with world > 1 exists
    obj, r_new, t_new, res, info = sba.bundle_adjust_points_and_extrinsics(
        p["points_2d"], pts0, pidx, cidx, K, D, p["R0"], p["t0"], None, return_info=True)
    ok = True
    if rank == 0:
        x0 = np.concatenate([np.concatenate([sba.rodrigues_to_vec(r) for r in p["R0"]]), p["t0"].ravel()])
        prob = sba.SBAProblem(p["points_2d"], pidx, cidx, K, D, n_pts, device=local)
        one = prob.solve(x0, pts0.astype(np.float64), max_nfev=1000, ftol=1e-10)
        dF = abs(info["cost"] - one["cost"]) / abs(one["cost"])
        dP = np.abs(obj - one["pts"]).max()
        dC = np.abs(info["params"] - one["params"]).max()
        dR = np.abs(res["after"] - one["fun"]).max()
        print(f"world {info['world']}: cost {info['cost']:.8e} vs single {one['cost']:.8e} (rel {dF:.2e}); nfev {info['nfev']} vs "
              f"{one['nfev']}; max|d pts| {dP:.2e} max|d cam| {dC:.2e} max|d res| {dR:.2e}; status {info['status']}/{one['status']}")
        ok = info["world"] == world and dF < 1e-8 and dP < 1e-5 and dC < 1e-6 and res["before"].shape == (2 * len(pidx),)
    flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
