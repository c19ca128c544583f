"""Per-phase cycle counts of bcr_factor (debug build with -DACINO_BCR_TIMING at scratch/libacino_bcr_timing.so):
block 0 / thread 0 timestamps after each barrier: load, diag0, (b) phases, (c) phases, write-out."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acinoset_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "..", "scratch", "libacino_bcr_timing.so")
L.lib = L._load()
import acinoset_b200 as ab, torch
from acinoset_b200 import lm
h = ab.Handle(0)
M = 3
rng = np.random.default_rng(0)
B = 75
W = rng.normal(0, 1, (M, B, B)); V = rng.normal(0, 0.5, (M, B, B))
D = np.stack([W[i].T @ W[i] + V[i].T @ V[i] + np.eye(B) for i in range(M)]); Lc = np.stack([W[i].T @ V[i - 1] * 0.3 for i in range(M)])
for i in range(1, M):
    D[i] += 0.1 * Lc[i] @ Lc[i].T; D[i - 1] += 0.1 * Lc[i].T @ Lc[i]
rhs = rng.normal(0, 1, (M, B))
dev = torch.device("cuda:0")
cs = lm.ChainSolver(h, M)
x = torch.zeros(M, B, dtype=torch.float64, device=dev)
for rep in range(3):
    Dd, Ld, rd = (torch.from_numpy(a).to(dev) for a in (D, Lc, rhs))
    torch.cuda.synchronize()
    L.lib.acino_debug_bcr_reset()
    lv = cs.levels[0]      # level 0: eliminate block 1 (two neighbours)
    h.call_dev("acino_bcr_factor_dev", lv["elim"].shape[0], lv["elim"], Dd, Ld, cs.P, cs.Q, cs.R, rd, cs.info)
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 8)()
    L.lib.acino_debug_bcr_cycles(out)
    cy = list(out)[:5]
    print("load %d  diag0 %d  (b) total %d  (c) total %d  write-out %d   sum %d cycles" % (cy[0], cy[1], cy[2], cy[3], cy[4], sum(cy)))
