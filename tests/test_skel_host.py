"""CPU: the generic-skeleton kernel bodies (acinoset_b200/csrc/skel_body.cuh), compiled for the host by
tests/host_harness/skel_host.cpp, against the NumPy oracle (oracle/skel_fte.py: build.py FK + complex-step Jacobian).
This checks the arithmetic the CUDA kernels run, in a container without a GPU; the GPU twin is tests/test_skel_gpu.py."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import skel_fte


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("skel_host") / "skel_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", str(out),
                           os.path.join(ROOT, "tests", "host_harness", "skel_host.cpp")])
    return ctypes.CDLL(str(out))


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def make_problem(tag, cams, n_frames, seed, noise=1.0):
    """States near the reference's shipped solution (K1/K2), measurements = projection + noise, a few dropped."""
    from acinoset_b200 import skeleton

    g = golden("generic_fk.npz")
    skel = json.loads(str(g[tag + "_skeleton_json"]))
    flat = skeleton.flatten_skeleton(skel)
    rng = np.random.default_rng(seed)
    K, D, R, t, _ = cams
    x = np.array(g[tag + "_x"][:n_frames], dtype=np.float64)
    x[:, 3:] += rng.normal(0, 0.2, x[:, 3:].shape)
    x[:, :3] = rng.uniform([1.0, 5.5, 0.8], [3.0, 7.5, 1.2], (n_frames, 3))       # inside the dummy scene
    f, names = skel_fte.pose_function(skel)
    assert names == flat["out_names"]
    P3 = np.array([f(r) for r in x])
    meas = np.stack([skel_fte.project(P3, K[c], D[c], R[c], t[c]) for c in range(len(K))], 1)
    meas = meas + rng.normal(0, noise, meas.shape)
    meas[rng.random(meas.shape[:-1]) < 0.05] += 300.0          # gross outliers
    w = np.where(rng.random(meas.shape[:-1]) < 0.85, 1.0 / skel_fte.MEAS_SIGMA_R, 0.0)
    x_eval = x + rng.normal(0, 0.02, x.shape)
    return skel, flat, x_eval, meas, w


def make_desc(lib, flat, cams, loss_kind, delta=0.05):
    K, D, R, t, _ = cams
    buf = np.zeros(lib.skel_host_desc_bytes() // 8 + 1, dtype=np.float64)
    n_parts, n_links, n_out = len(flat["parts"]), len(flat["link_parent"]), len(flat["out_order"])
    Kc, Dc, Rc, tc = (np.ascontiguousarray(a, dtype=np.float64) for a in (K, D.reshape(-1, 4), R, t.reshape(-1, 3)))
    lib.skel_host_make_desc(_ptr(buf), n_parts, n_links, n_out, len(K), _ptr(flat["dof_mask"]), _ptr(flat["link_parent"]),
                            _ptr(flat["link_flags"]), _ptr(flat["link_tv"]), _ptr(flat["out_path"]), loss_kind,
                            ctypes.c_double(3.0), ctypes.c_double(10.0), ctypes.c_double(20.0), ctypes.c_double(delta),
                            _ptr(Kc), _ptr(Dc), _ptr(Rc), _ptr(tc))
    return buf


@pytest.mark.parametrize("tag,loss", [("K1", "abs"), ("K1", "redescending"), ("K2", "abs")])
def test_skel_eval_body_matches_oracle(harness, dummy_cams, tag, loss):
    skel, flat, x, meas, w = make_problem(tag, dummy_cams, 5, seed=3)
    K, D, R, t, _ = dummy_cams
    N, P = x.shape
    desc = make_desc(harness, flat, dummy_cams, 0 if loss == "redescending" else 1)
    cost = np.zeros(N)
    g = np.zeros((N, P))
    Hu = np.zeros((N, P * (P + 1) // 2))
    harness.skel_host_eval(_ptr(desc), N, _ptr(x), _ptr(np.ascontiguousarray(meas)), _ptr(np.ascontiguousarray(w)),
                           _ptr(cost), _ptr(g), _ptr(Hu))
    c0, g0, H0 = skel_fte.skel_eval(skel, x, meas, w, K, D, R, t, loss=loss)
    assert np.abs(cost - c0).max() < 1e-9 * np.abs(c0).max()
    assert np.abs(g - g0).max() < 1e-9 * np.abs(g0).max()
    H = skel_fte.upper_unpack(Hu, P)
    assert np.abs(H - H0).max() < 1e-9 * np.abs(H0).max()
    # unused slots (leaf parts, switched-off dofs) have exactly zero rows
    unused = np.abs(H0).sum(axis=(0, 1)) == 0
    assert unused.sum() > 0 and np.all(H[:, unused] == 0) and np.all(g[:, unused] == 0)


def test_lm_building_blocks_match_dense_algebra(harness, dummy_cams):
    """prepare / assemble / band Cholesky / trial / pred against dense NumPy on a small problem."""
    skel, flat, x, meas, w = make_problem("K1", dummy_cams, 9, seed=5)
    K, D, R, t, _ = dummy_cams
    N, P = x.shape
    L = (P - 3) // 3
    rng = np.random.default_rng(0)
    c0, g0, H0 = skel_fte.skel_eval(skel, x, meas, w, K, D, R, t, loss="abs")
    Hu = np.ascontiguousarray(skel_fte.upper_pack(H0))
    sw = np.full(P, 2 * skel_fte.MODEL_WEIGHT / (1 / 120.0) ** 4) * rng.uniform(0.5, 1.5, P) * 1e-6
    lo, hi = skel_fte.bounds(L)
    x = np.clip(x, lo, hi)
    x[2, 5] = hi[5]          # one variable sitting on its bound
    x[N - 1, 6] = 3.0        # beyond the bound in the last (free) frame
    gtot = np.zeros((N, P))
    fixed = np.zeros((N, P), dtype=np.uint8)
    cs = np.zeros(N)
    harness.skel_host_prepare(N, P, 1, _ptr(x), _ptr(g0), _ptr(sw), _ptr(lo), _ptr(hi), _ptr(gtot), _ptr(fixed), _ptr(cs))
    G3 = skel_fte.d3_matrix(N).T @ skel_fte.d3_matrix(N)
    gt_ref = g0 + (G3 @ x) * sw
    assert np.abs(gtot - gt_ref).max() < 1e-9 * np.abs(gt_ref).max()
    assert abs(cs.sum() - skel_fte.smooth_cost(x, sw)) < 1e-9 * max(1.0, cs.sum())
    lo_f, hi_f = np.tile(lo, (N, 1)), np.tile(hi, (N, 1))
    lo_f[-1], hi_f[-1] = -np.inf, np.inf
    fx_ref = ((x <= lo_f) & (gt_ref > 0)) | ((x >= hi_f) & (gt_ref < 0))
    assert np.array_equal(fixed.astype(bool), fx_ref)
    lam = 0.37
    hb = 3 * P
    AB = np.zeros((N * P, hb + 1))
    rhs = np.zeros(N * P)
    harness.skel_host_assemble(N, P, _ptr(Hu), _ptr(gtot), _ptr(fixed), _ptr(sw), ctypes.c_double(lam), _ptr(AB), _ptr(rhs))
    B = np.kron(G3, np.diag(sw))
    for n in range(N):
        B[n * P:(n + 1) * P, n * P:(n + 1) * P] += H0[n]
    Bd = B + lam * np.diag(np.diag(B))
    f = fx_ref.ravel()
    Bd[f, :] = 0
    Bd[:, f] = 0
    Bd[f, f] = 1
    dense = np.zeros_like(Bd)
    for i in range(N * P):
        for k in range(min(i, hb) + 1):
            dense[i, i - k] = dense[i - k, i] = AB[i, k]
    assert np.abs(dense - Bd).max() < 1e-12 * np.abs(Bd).max()
    d_ref = np.linalg.solve(Bd, np.where(f, 0.0, -gt_ref.ravel()))
    info = np.zeros(1, dtype=np.int32)
    for nb in (16, 7, 1):                     # the kernel's panel width, a ragged one, and the unblocked limit
        ABc, xs = AB.copy(), rhs.copy()
        info[:] = 0
        harness.skel_host_band_solve(ctypes.c_longlong(N * P), hb, nb, _ptr(ABc), _ptr(xs), _ptr(info))
        assert info[0] == 0
        assert np.abs(xs - d_ref).max() < 1e-8 * np.abs(d_ref).max()
    Lf = np.linalg.cholesky(Bd)               # the band holds the Cholesky factor itself
    for i in range(0, N * P, 37):
        for k in range(min(i, hb) + 1):
            assert abs(ABc[i, k] - Lf[i, i - k]) < 1e-9 * max(1.0, abs(Lf[i, i - k]))
    rhs = xs
    xt = np.zeros_like(x)
    d = rhs.reshape(N, P).copy()
    harness.skel_host_trial(N, P, 1, _ptr(x), _ptr(d), _ptr(lo), _ptr(hi), _ptr(xt))
    assert np.array_equal(xt, np.clip(x + d, lo_f, hi_f))
    pred = np.zeros(N)
    step = np.zeros(N)
    harness.skel_host_pred(N, P, _ptr(x), _ptr(xt), _ptr(gtot), _ptr(Hu), _ptr(sw), _ptr(pred), _ptr(step))
    s = (xt - x).ravel()
    pred_ref = -(gt_ref.ravel() @ s) - 0.5 * s @ (B @ s)
    assert abs(pred.sum() - pred_ref) < 1e-9 * abs(pred_ref)
    assert step.max() == np.abs(s).max()
    # a non-positive pivot is reported, not silently factored
    bad = np.array([[1.0, 0.0], [2.0, 1.0]])     # rows [B00], [B11, B10]: B = [[1,1],[1,... ]] -> second pivot 2 - 1 = 1 ok
    bad = np.array([[1.0, 0.0], [0.5, 2.0]])     # B = [[1,2],[2,0.5]]: second pivot 0.5 - 4 < 0
    xb = np.ones(2)
    info[:] = 0
    harness.skel_host_band_solve(ctypes.c_longlong(2), 1, 16, _ptr(bad), _ptr(xb), _ptr(info))
    assert info[0] == 2


def test_band_cholesky_random_systems(harness):
    """Blocked band Cholesky + substitution on random SPD band systems, sizes that are not multiples of the panel."""
    rng = np.random.default_rng(4)
    for n, hb, nb in [(1, 0, 16), (5, 2, 16), (50, 7, 16), (131, 20, 16), (200, 45, 8), (64, 63, 16), (97, 30, 32)]:
        A = np.zeros((n, n))
        for i in range(n):
            for k in range(1, min(i, hb) + 1):
                A[i, i - k] = A[i - k, i] = rng.normal()
        A += np.diag(np.abs(A).sum(1) + rng.uniform(0.5, 2.0, n))
        AB = np.zeros((n, hb + 1))
        for i in range(n):
            for k in range(min(i, hb) + 1):
                AB[i, k] = A[i, i - k]
        b = rng.normal(size=n)
        x = b.copy()
        info = np.zeros(1, dtype=np.int32)
        harness.skel_host_band_solve(ctypes.c_longlong(n), hb, nb, _ptr(AB), _ptr(x), _ptr(info))
        assert info[0] == 0
        ref = np.linalg.solve(A, b)
        assert np.abs(x - ref).max() < 1e-10 * max(1.0, np.abs(ref).max()), (n, hb, nb)


def test_kernel_thread_mappings_with_real_threads(harness, dummy_cams):
    """The same bodies on T real threads with real barriers (a CTA in slow motion): the strided / 2-D thread mappings and
    the participants-only barrier of the panel step give the one-thread results."""
    rng = np.random.default_rng(9)
    for n, hb, nb, T in [(131, 20, 16, 64), (200, 45, 8, 128), (97, 30, 16, 48), (40, 39, 8, 128)]:
        A = np.zeros((n, n))
        for i in range(n):
            for k in range(1, min(i, hb) + 1):
                A[i, i - k] = A[i - k, i] = rng.normal()
        A += np.diag(np.abs(A).sum(1) + rng.uniform(0.5, 2.0, n))
        AB = np.zeros((n, hb + 1))
        for i in range(n):
            for k in range(min(i, hb) + 1):
                AB[i, k] = A[i, i - k]
        b = rng.normal(size=n)
        AB1, x1, AB2, x2 = AB.copy(), b.copy(), AB.copy(), b.copy()
        info = np.zeros(1, dtype=np.int32)
        harness.skel_host_band_solve(ctypes.c_longlong(n), hb, nb, _ptr(AB1), _ptr(x1), _ptr(info))
        assert harness.skel_host_band_solve_mt(ctypes.c_longlong(n), hb, nb, _ptr(AB2), _ptr(x2), _ptr(info), T) == 0
        assert info[0] == 0
        assert np.array_equal(x1, x2) and np.array_equal(AB1, AB2), (n, hb, nb, T)      # same operations, same order per entry
        assert np.abs(x2 - np.linalg.solve(A, b)).max() < 1e-10
    # a non-positive pivot stops every thread (participants and bystanders) without a deadlock
    bad = np.zeros((40, 9))
    bad[:, 0] = 1.0
    bad[17, 0] = -1.0
    xb = np.ones(40)
    info[:] = 0
    assert harness.skel_host_band_solve_mt(ctypes.c_longlong(40), 8, 8, _ptr(bad), _ptr(xb), _ptr(info), 128) == 0
    assert info[0] == 18
    # skel_eval with 96 real threads per frame
    skel, flat, x, meas, w = make_problem("K1", dummy_cams, 2, seed=13)
    N, P = x.shape
    desc = make_desc(harness, flat, dummy_cams, 1)
    out = []
    for mt in (0, 96):
        cost, g, Hu = np.zeros(N), np.zeros((N, P)), np.zeros((N, P * (P + 1) // 2))
        args = (_ptr(desc), N, _ptr(x), _ptr(np.ascontiguousarray(meas)), _ptr(np.ascontiguousarray(w)), _ptr(cost), _ptr(g), _ptr(Hu))
        harness.skel_host_eval_mt(*args, mt) if mt else harness.skel_host_eval(*args)
        out.append((cost, g, Hu))
    for a, b2 in zip(out[0], out[1]):
        assert np.array_equal(a, b2)
