"""Worker for tests/test_multi_rank_cpu.py (world_size-2 gloo on CPU): exercises the product's
multi-rank host logic (sharding, interface packing, all_gather, chain assembly, scatter) with the
NumPy oracle standing in for the CUDA block kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acinoset_b200 import bcr, lm
    from oracle import bcr as obcr

    B = lm.SB
    Mloc = [5, 4][rank] if world == 2 else 4
    Ms = [5, 4] if world == 2 else [4] * world
    M = sum(Ms)
    rng = np.random.default_rng(1234)          # same global chain on every rank
    D = np.zeros((M, B, B))
    Lc = np.zeros((M, B, B))
    W = rng.normal(0, 1, (M, B, B))
    V = rng.normal(0, 0.5, (M, B, B))
    for i in range(M):
        D[i] = W[i].T @ W[i] + V[i].T @ V[i] + np.eye(B)
        if i > 0:
            Lc[i] = W[i].T @ V[i - 1] * 0.3
            D[i] += 0.1 * Lc[i] @ Lc[i].T
            D[i - 1] += 0.1 * Lc[i].T @ Lc[i]
    rhs = rng.normal(0, 1, (M, B))
    # rank 0 freezes two variables of its last block: rows/cols zeroed locally, and the columns of
    # rank 1's first coupling block must be zeroed by gather_interface_chain
    start = sum(Ms[:rank])
    fixed = np.zeros((M, B), dtype=bool)
    fixed[Ms[0] - 1, [3, 40]] = True
    Dg, Lg, rg = D.copy(), Lc.copy(), rhs.copy()
    for (i, k) in zip(*np.nonzero(fixed)):
        Dg[i][k, :] = 0
        Dg[i][:, k] = 0
        Dg[i][k, k] = 1
        Lg[i][k, :] = 0
        if i + 1 < M:
            Lg[i + 1][:, k] = 0
        rg[i, k] = 0
    x_ref = np.linalg.solve(obcr.dense_from_chain(Dg, Lg), rg.ravel()).reshape(M, B)

    # what a rank can build on its own: its blocks with ITS OWN frozen variables applied, but the
    # first coupling block only knows the band (not the neighbour's frozen set)
    sl = slice(start, start + Mloc)
    Dl, Ll, rl = Dg[sl].copy(), Lg[sl].copy(), rg[sl].copy()
    if rank > 0:
        Ll[0] = Lc[start].copy()
        own = fixed[start]
        Ll[0][own, :] = 0
    levels, left = bcr.make_schedule(Mloc, True, True)
    assert left == [0, Mloc - 1]
    fac = obcr.bcr_reduce(Dl, Ll, rl, levels)
    payload = lm.pack_interface(torch.from_numpy(Dl), torch.from_numpy(Ll), torch.from_numpy(rl),
                                torch.from_numpy(fixed[sl].astype(np.uint8)))
    assert payload.numel() == lm.PAYLOAD
    Dc, Lcc, rc = lm.gather_interface_chain(payload, world)
    clevels, cleft = bcr.make_schedule(2 * world)
    xc = obcr.bcr_solve(Dc.numpy(), Lcc.numpy(), rc.numpy(), clevels)
    xf, xl, hl, hr = lm.split_interface_solution(torch.from_numpy(xc), rank, world)
    x = np.zeros((Mloc, B))
    x[0], x[Mloc - 1] = xf.numpy(), xl.numpy()
    obcr.bcr_backsub(Dl, rl, fac, levels, x)
    err = np.abs(x - x_ref[sl]).max()
    ok = err < 1e-8 * max(1.0, np.abs(x_ref).max())
    if rank > 0:
        ok = ok and np.abs(hl.numpy() - x_ref[start - 1]).max() < 1e-8
    if rank < world - 1:
        ok = ok and np.abs(hr.numpy() - x_ref[start + Mloc]).max() < 1e-8
    res = torch.tensor([1.0 if ok else 0.0, err])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    print(f"rank {rank} err {err:.3e} ok {ok}")
    sys.exit(0 if res[0].item() == 1.0 else 1)


if __name__ == "__main__":
    main()
