"""Oracle (test infrastructure): NumPy fp64 executor of the block-cyclic-reduction schedule used by
the CUDA LM solver, plus a dense reference solve.  Checks the algebra of acinoset_b200/bcr.py +
csrc/bcr.cu (the reference itself has no such solver - it shells out to IPOPT,
/root/reference/src/all_optimizations.py:503-524)."""
import numpy as np


def dense_from_chain(D, Lc):
    M, B, _ = D.shape
    A = np.zeros((M * B, M * B))
    for i in range(M):
        A[i * B:(i + 1) * B, i * B:(i + 1) * B] = D[i]
        if i > 0:
            A[i * B:(i + 1) * B, (i - 1) * B:i * B] = Lc[i]
            A[(i - 1) * B:i * B, i * B:(i + 1) * B] = Lc[i].T
    return A


def bcr_reduce(D, Lc, rhs, levels):
    """In-place forward phase.  Returns the factors dict(R, P, Q) (z is left in rhs)."""
    M, B, _ = D.shape
    P = np.zeros_like(D)
    Q = np.zeros_like(D)
    for lv in levels:
        for e, a, c in lv["elim"]:
            R = np.linalg.cholesky(D[e])
            D[e] = R
            if a >= 0:
                P[e] = np.linalg.solve(R, Lc[e])
            if c >= 0:
                Q[e] = np.linalg.solve(R, Lc[c].T)
            rhs[e] = np.linalg.solve(R, rhs[e])
        for j, el, er in lv["surv"]:
            if el >= 0:      # j was the right neighbour of el
                D[j] -= Q[el].T @ Q[el]
                rhs[j] -= Q[el].T @ rhs[el]
                Lc[j] = -Q[el].T @ P[el]
            if er >= 0:      # j was the left neighbour of er
                D[j] -= P[er].T @ P[er]
                rhs[j] -= P[er].T @ rhs[er]
    return dict(P=P, Q=Q)


def bcr_backsub(D, rhs, fac, levels, x):
    """x (M,B): entries of the blocks that were never eliminated must be filled in by the caller."""
    for lv in reversed(levels):
        for e, a, c in lv["elim"]:
            v = rhs[e].copy()
            if a >= 0:
                v -= fac["P"][e] @ x[a]
            if c >= 0:
                v -= fac["Q"][e] @ x[c]
            x[e] = np.linalg.solve(D[e].T, v)
    return x


def bcr_solve(D, Lc, rhs, levels):
    D, Lc, rhs = D.copy(), Lc.copy(), rhs.copy()
    fac = bcr_reduce(D, Lc, rhs, levels)
    x = np.zeros_like(rhs)
    return bcr_backsub(D, rhs, fac, levels, x)
