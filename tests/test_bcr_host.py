"""CPU: the BCR elimination schedule (product host code) executed by the NumPy oracle equals a
dense solve; pinned ends give the Schur complement onto the interface blocks."""
import numpy as np
import pytest

from acinoset_b200 import bcr
from oracle import bcr as obcr


def _random_chain(M, B, rng):
    # SPD block tridiagonal: A = G^T G + I with G banded
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
    return D, Lc, rhs


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 7, 8, 16, 33])
def test_bcr_matches_dense(M):
    rng = np.random.default_rng(M)
    D, Lc, rhs = _random_chain(M, 6, rng)
    levels, left = bcr.make_schedule(M)
    assert left == []
    x = obcr.bcr_solve(D, Lc, rhs, levels)
    xd = np.linalg.solve(obcr.dense_from_chain(D, Lc), rhs.ravel()).reshape(M, -1)
    assert np.abs(x - xd).max() < 1e-9 * max(1, np.abs(xd).max())
    # every block eliminated exactly once
    el = np.concatenate([lv["elim"][:, 0] for lv in levels])
    assert sorted(el.tolist()) == list(range(M))


@pytest.mark.parametrize("M", [2, 3, 4, 9, 20])
def test_bcr_pinned_gives_schur_complement(M):
    rng = np.random.default_rng(100 + M)
    B = 5
    D, Lc, rhs = _random_chain(M, B, rng)
    A = obcr.dense_from_chain(D, Lc)
    levels, left = bcr.make_schedule(M, pin_first=True, pin_last=True)
    assert left == [0, M - 1]
    D2, Lc2, r2 = D.copy(), Lc.copy(), rhs.copy()
    fac = obcr.bcr_reduce(D2, Lc2, r2, levels)
    # dense Schur complement onto blocks {0, M-1}
    idx_s = np.r_[0:B, (M - 1) * B:M * B]
    idx_i = np.setdiff1d(np.arange(M * B), idx_s)
    Aii, Ais, Ass = A[np.ix_(idx_i, idx_i)], A[np.ix_(idx_i, idx_s)], A[np.ix_(idx_s, idx_s)]
    S = Ass - Ais.T @ np.linalg.solve(Aii, Ais) if idx_i.size else Ass
    rs = rhs.ravel()[idx_s] - (Ais.T @ np.linalg.solve(Aii, rhs.ravel()[idx_i]) if idx_i.size else 0)
    S_bcr = np.block([[D2[0], Lc2[M - 1].T], [Lc2[M - 1], D2[M - 1]]])
    assert np.abs(S_bcr - S).max() < 1e-9 * np.abs(S).max()
    assert np.abs(np.r_[r2[0], r2[M - 1]] - rs).max() < 1e-9 * max(1, np.abs(rs).max())
    # interface solve + back-substitution reproduces the dense solution
    xs = np.linalg.solve(S, rs)
    x = np.zeros((M, B))
    x[0], x[M - 1] = xs[:B], xs[B:]
    obcr.bcr_backsub(D2, r2, fac, levels, x)
    xd = np.linalg.solve(A, rhs.ravel()).reshape(M, B)
    assert np.abs(x - xd).max() < 1e-9 * max(1, np.abs(xd).max())
