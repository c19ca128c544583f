"""Worker for tests/test_multi_rank_cpu.py (world_size-2 gloo on CPU): the sharded SBA host logic
(point sharding, local Schur complement -> ONE all_reduce of [S | rhs] -> redundant camera solve ->
scatter of the point results) with the NumPy oracle standing in for the CUDA kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def normal_blocks(x, C, n_pts, pidx, cidx, K, D, p2d):
    """Per-observation blocks in the CUDA layout (Jc = [d/d rvec | d/d t]) from the fp64 oracle."""
    from oracle import sba as osba

    Jr, Jt, Jx = osba.jac_blocks_points_extrinsics(x, C, n_pts, pidx, cidx, K, D)
    f = osba.cost_func_points_extrinsics(x, C, n_pts, pidx, cidx, K, D, p2d).reshape(-1, 2)
    w = 1.0 / (1.0 + f ** 2)                       # Cauchy IRLS weight, f_scale = 1
    return f, np.concatenate([Jr, Jt], axis=2), Jx, w


def reduced_system(f, Jc, Jx, w, pidx, cidx, C, n_pts, lam):
    """Schur complement of the point blocks onto the 6C camera system, [S | rhs] packed."""
    n6 = 6 * C
    U = np.zeros((n6, n6))
    g = np.zeros(n6)
    V = np.zeros((n_pts, 3, 3))
    gv = np.zeros((n_pts, 3))
    W = np.zeros((n_pts, n6, 3))
    for i in range(len(pidx)):
        a, p = cidx[i], pidx[i]
        sl = slice(6 * a, 6 * a + 6)
        U[sl, sl] += Jc[i].T @ (w[i][:, None] * Jc[i])
        g[sl] += Jc[i].T @ (w[i] * f[i])
        V[p] += Jx[i].T @ (w[i][:, None] * Jx[i])
        gv[p] += Jx[i].T @ (w[i] * f[i])
        W[p][sl] += Jc[i].T @ (w[i][:, None] * Jx[i])
    S = U + lam * np.diag(np.diag(U))
    rhs = -g
    Vi = np.empty_like(V)
    for p in range(n_pts):
        Vd = V[p] + lam * np.diag(np.diag(V[p]))
        Vi[p] = np.linalg.inv(Vd)
        S -= W[p] @ Vi[p] @ W[p].T
        rhs += W[p] @ Vi[p] @ gv[p]
    return S, rhs, Vi, gv, W


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from acinoset_b200 import sba
    from oracle import fisheye

    p = synth.make_sba_problem(24, fisheye.project, seed=11)          # same problem on every rank
    K, D = p["K"], p["D"]
    C = len(K)
    pidx, cidx, p2d = p["point_3d_indices"], p["camera_indices"], p["points_2d"].astype(np.float64)
    n_pts = len(p["points_3d_true"])
    rv = np.concatenate([fisheye.rodrigues_inv(r) for r in p["R0"]])
    pts0 = p["points_3d_true"] + np.random.default_rng(5).normal(0, 0.01, (n_pts, 3))
    x = np.concatenate([rv, p["t0"].ravel(), pts0.ravel()])
    lam = 1e-3
    # global reference step (what a single rank would compute)
    f, Jc, Jx, w = normal_blocks(x, C, n_pts, pidx, cidx, K, D, p2d)
    S, rhs, Vi, gv, W = reduced_system(f, Jc, Jx, w, pidx, cidx, C, n_pts, lam)
    dc_ref = np.linalg.solve(S, rhs)
    dp_ref = np.stack([-Vi[q] @ (gv[q] + W[q].T @ dc_ref) for q in range(n_pts)])
    # sharded: my points only
    plan = sba.shard_points(pidx, n_pts, world)
    assert sum(n for _, n, _ in plan) == n_pts and sum(len(i) for _, _, i in plan) == len(pidx)
    assert len(np.unique(np.concatenate([i for _, _, i in plan]))) == len(pidx)
    p0, npl, ids = plan[rank]
    assert np.all((pidx[ids] >= p0) & (pidx[ids] < p0 + npl))
    Sl, rl, Vil, gvl, Wl = reduced_system(f[ids], Jc[ids], Jx[ids], w[ids], pidx[ids] - p0, cidx[ids], C, npl, lam)
    n6 = 6 * C
    Sr = torch.from_numpy(np.concatenate([Sl.ravel(), rl]))
    sba.allreduce_camera_system(Sr, world)
    Sg, rg = Sr[:n6 * n6].view(n6, n6).numpy(), Sr[n6 * n6:].numpy()
    dc = np.linalg.solve(Sg, rg)
    dp = np.stack([-Vil[q] @ (gvl[q] + Wl[q].T @ dc) for q in range(npl)])
    full = sba.scatter_sum(torch.from_numpy(dp), torch.arange(p0, p0 + npl), n_pts, world).numpy()
    e1 = np.abs(dc - dc_ref).max() / np.abs(dc_ref).max()
    e2 = np.abs(full - dp_ref).max() / np.abs(dp_ref).max()
    sums = sba.allreduce_scalars(torch.tensor([float((w[ids] * f[ids] ** 2).sum())], dtype=torch.float64), world)
    e3 = abs(sums.item() - float((w * f ** 2).sum())) / float((w * f ** 2).sum())
    ok = e1 < 1e-9 and e2 < 1e-9 and e3 < 1e-12
    res = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    print(f"rank {rank} camera step err {e1:.2e} point step err {e2:.2e} sum err {e3:.2e} ok {ok}")
    sys.exit(0 if res[0].item() == 1.0 else 1)


if __name__ == "__main__":
    main()
