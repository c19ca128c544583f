"""Block cyclic reduction (BCR) of the block-tridiagonal normal equations of the FTE
Levenberg-Marquardt step - host side: elimination schedule and the driver of the CUDA kernels.

The LM system  (blockdiag(H_n) + S + lam D) dx = -g  is block-heptadiagonal in 25x25 frame
blocks (the smoothness term of all_optimizations.py:369-391 couples frames n-3..n+3 and is
diagonal across parameters).  Grouping 3 frames per super-block (75x75) makes it block
TRIdiagonal; cyclic reduction eliminates every other super-block per level (all eliminations
of a level are independent -> one CTA each), log2(M) levels, then back-substitutes in reverse.
With the two end super-blocks "pinned" the same code produces the rank-local Schur complement
onto its interface blocks for the multi-GPU solve (one all_gather of 2 blocks per rank).
"""
import numpy as np

SB_FRAMES = 3
SB = 75  # super-block size = 3 frames x 25 parameters


def make_schedule(M, pin_first=False, pin_last=False):
    """Elimination schedule for a chain of M super-blocks.

    Returns a list of levels; each level is a dict of int32 arrays
      elim  (ne,3): (block e, left neighbour a or -1, right neighbour c or -1)
      surv  (ns,3): (block j, eliminated left neighbour or -1, eliminated right neighbour or -1)
    Blocks 0 / M-1 are never eliminated when pinned.  Without pins the last level eliminates the
    single remaining block (no neighbours)."""
    active = list(range(M))
    levels = []
    while True:
        n = len(active)
        pins = set()
        if pin_first:
            pins.add(active[0])
        if pin_last:
            pins.add(active[-1])
        if n == 0 or all(b in pins for b in active):
            break
        if n == 1:
            elim_pos = [0]
        else:
            elim_pos = [k for k in range(1, n, 2) if active[k] not in pins]
            if not elim_pos:  # e.g. [pinned, free] or [free, pinned]: eliminate the free one
                elim_pos = [k for k in range(n) if active[k] not in pins][:1]
        eset = set(elim_pos)
        elim, surv = [], []
        for k in elim_pos:
            a = active[k - 1] if k - 1 >= 0 else -1
            c = active[k + 1] if k + 1 < n else -1
            elim.append((active[k], a, c))
        for k in range(n):
            if k in eset:
                continue
            el = active[k - 1] if (k - 1) in eset else -1
            er = active[k + 1] if (k + 1) in eset else -1
            if el >= 0 or er >= 0:
                surv.append((active[k], el, er))
        levels.append(dict(elim=np.array(elim, dtype=np.int32).reshape(-1, 3),
                           surv=np.array(surv, dtype=np.int32).reshape(-1, 3)))
        active = [active[k] for k in range(n) if k not in eset]
    return levels, active
