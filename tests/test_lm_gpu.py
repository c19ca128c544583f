"""GPU: BCR chain solver vs dense solve; LM building blocks vs the oracle; full FTE solve vs the
fp64 CPU restatement of the same algorithm (oracle/lm.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle(dummy_cams):
    import acinoset_b200 as ab

    K, D, R, t, _ = dummy_cams
    h = ab.Handle(0)
    h.set_cameras(K, D, R, t)
    yield h
    h.close()


def _spd_chain(M, rng, B=75):
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
    return D, Lc, rng.normal(0, 1, (M, B))


@pytest.mark.parametrize("M", [1, 2, 3, 5, 8, 33, 100])
def test_bcr_cuda_matches_dense(handle, M):
    import torch
    from acinoset_b200 import lm
    from oracle import bcr as obcr

    rng = np.random.default_rng(M)
    D, Lc, rhs = _spd_chain(M, rng)
    xd = np.linalg.solve(obcr.dense_from_chain(D, Lc), rhs.ravel()).reshape(M, -1)
    dev = torch.device("cuda:0")
    Dd, Ld, rd = (torch.from_numpy(a).to(dev) for a in (D, Lc, rhs))
    x = torch.zeros(M, 75, dtype=torch.float64, device=dev)
    cs = lm.ChainSolver(handle, M)
    cs.solve(Dd, Ld, rd, x)
    torch.cuda.synchronize()
    assert int(cs.info.item()) == 0
    assert np.abs(x.cpu().numpy() - xd).max() < 1e-9 * max(1.0, np.abs(xd).max())


def test_bcr_cuda_pinned_matches_oracle(handle):
    import torch
    from acinoset_b200 import bcr, lm
    from oracle import bcr as obcr

    M = 21
    rng = np.random.default_rng(7)
    D, Lc, rhs = _spd_chain(M, rng)
    levels, _ = bcr.make_schedule(M, True, True)
    D2, L2, r2 = D.copy(), Lc.copy(), rhs.copy()
    obcr.bcr_reduce(D2, L2, r2, levels)
    dev = torch.device("cuda:0")
    Dd, Ld, rd = (torch.from_numpy(a).to(dev) for a in (D, Lc, rhs))
    cs = lm.ChainSolver(handle, M, pinned=True)
    cs.reduce(Dd, Ld, rd)
    torch.cuda.synchronize()
    for a, b in ((Dd[0], D2[0]), (Dd[M - 1], D2[M - 1]), (Ld[M - 1], L2[M - 1]), (rd[0], r2[0]), (rd[M - 1], r2[M - 1])):
        assert np.abs(a.cpu().numpy() - b).max() < 1e-9 * max(1.0, np.abs(b).max())


def test_lm_assemble_matches_oracle_system(handle, dummy_cams):
    """prepare + assemble reproduce B + lam diag(B) and -g of the oracle (dense comparison)."""
    import synth
    import torch
    from acinoset_b200 import lm
    from oracle import bcr as obcr, fisheye, fte, lm as olm, skeleton

    N = 31
    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=21, cams=dummy_cams)
    K, D, R, t, _ = dummy_cams
    x0 = p["x0"].copy()
    lo, hi = skeleton.active_bounds()
    for pp in range(12, 20):  # whole columns on a bound (no smoothness gradient): some get frozen
        x0[:, pp] = hi[pp] if pp % 2 else lo[pp]
    x0 = np.clip(x0, lo, hi)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"])
    s = sol.st[0]
    s["x_ext"][3:3 + N] = torch.from_numpy(x0).cuda()
    s["x32"].copy_(s["x_ext"][3:3 + N].float())
    sol._eval(s)
    sol._prepare(s)
    lam = 0.37
    sol.h.call_dev("acino_lm_assemble_dev", N, 0, N, sol.M, s["H"], s["gtot"], s["fixed"], sol.sw, lam, sol.D, sol.Lc, sol.rhs)
    torch.cuda.synchronize()
    A = obcr.dense_from_chain(sol.D.cpu().numpy(), sol.Lc.cpu().numpy())[:N * 25, :N * 25]
    rhs = sol.rhs.cpu().numpy().ravel()[:N * 25]
    # oracle system from the SAME fp32-rounded evaluation inputs
    x32 = x0.astype(np.float32).astype(np.float64)
    c, g, H = fte.fte_eval(x32, p["meas"].astype(np.float32).astype(np.float64), p["w"].astype(np.float32).astype(np.float64), K, D, R, t)
    q = fte.model_weights_active()
    gt = (g + fte.smooth_grad(x0, p["Ts"], q)).ravel()
    B = (olm.assemble(H, N) + olm.smooth_matrix(N, p["Ts"], q)).toarray()
    xv = x0.ravel()
    fixed = ((xv <= np.tile(lo, N)) & (gt > 0)) | ((xv >= np.tile(hi, N)) & (gt < 0))
    assert fixed.sum() >= 1
    assert np.array_equal(s["fixed"].cpu().numpy().ravel().astype(bool), fixed)
    Aref = B + lam * np.diag(np.diag(B))
    Aref[fixed, :] = 0
    Aref[:, fixed] = 0
    Aref[fixed, fixed] = 1
    rref = -gt.copy()
    rref[fixed] = 0
    scale = np.sqrt(np.outer(np.diag(Aref), np.diag(Aref)))
    assert (np.abs(A - Aref) / scale).max() < 2e-5       # data blocks come from the fp32 kernel
    assert np.abs(rhs - rref).max() < 2e-4 * np.abs(rref).max()
    # smoothness cost
    cs = float(s["cost_s"].sum().item())
    assert abs(cs - fte.smooth_cost(x0, p["Ts"], q)) < 1e-9 * max(1.0, cs)


def _state_on_bounds(sol, p, N, seed=0):
    """Put an iterate with some whole columns on their bounds into the accepted state and evaluate it."""
    import torch
    from oracle import skeleton

    lo, hi = skeleton.active_bounds()
    x0 = p["x0"].copy()
    for pp in range(12, 20):
        x0[:, pp] = hi[pp] if pp % 2 else lo[pp]
    x0 = np.clip(x0, lo, hi)
    s = sol.st[0]
    s["x_ext"].zero_()
    s["x_ext"][3:3 + N] = torch.from_numpy(x0).cuda()
    s["x32"].copy_(s["x_ext"][3:3 + N].float())
    sol._eval(s)
    sol._prepare(s)
    return s


@pytest.mark.parametrize("N", [1, 2, 3, 4, 6, 7, 31, 100, 302])
def test_structured_level0_matches_dense_route(handle, dummy_cams, N):
    """REDUCE + BACKSUB of the plan (structured level 0 fused with the assembly, csrc/lm_l0.cu) give the same step as
    lm_assemble + the dense block cyclic reduction on every level."""
    import synth
    import torch
    from acinoset_b200 import lm
    from oracle import fisheye, skeleton

    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=31 + N, cams=dummy_cams)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"])
    s = _state_on_bounds(sol, p, N)
    lam = 0.05
    D, Lc, rhs = (torch.zeros_like(t) for t in (sol.D, sol.Lc, sol.rhs))
    sol.h.call_dev("acino_lm_assemble_dev", N, 0, N, sol.M, s["H"], s["gtot"], s["fixed"], sol.sw, lam, D, Lc, rhs)
    x_ref = torch.zeros_like(rhs)
    cs = lm.ChainSolver(handle, sol.M)
    cs.solve(D, Lc, rhs, x_ref)
    dx, dhalo = sol.linear_solve(lam)
    torch.cuda.synchronize()
    assert int(cs.info.item()) == 0 and int(sol.info.item()) == 0
    xr, xn = x_ref.cpu().numpy(), dx.cpu().numpy()
    assert np.abs(xr).max() > 1e-6
    assert np.abs(xn - xr).max() < 1e-8 * np.abs(xr).max()
    assert np.abs(dhalo.cpu().numpy()).max() == 0.0


@pytest.mark.parametrize("N,frame0,ng", [(6, 0, 1000), (150, 300, 1000), (151, 849, 1000), (9, 3, 12)])
def test_structured_level0_pinned_payload_matches_dense_route(handle, dummy_cams, N, frame0, ng):
    """A rank's shard (pinned end blocks, global frame offset): the interface payload of the plan equals the
    end blocks of lm_assemble + pinned dense reduction."""
    import synth
    import torch
    from acinoset_b200 import lm
    from oracle import fisheye, skeleton

    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=77 + N, cams=dummy_cams, start=frame0)
    last = frame0 + N == ng
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"], frame0=frame0, n_global=ng, rank=1 if frame0 else 0,
                       world=3 if not last else 2)
    s = _state_on_bounds(sol, p, N)
    lam = 0.2
    M = sol.M
    D, Lc, rhs = (torch.zeros_like(t) for t in (sol.D, sol.Lc, sol.rhs))
    sol.h.call_dev("acino_lm_assemble_dev", N, frame0, ng, M, s["H"], s["gtot"], s["fixed"], sol.sw, lam, D, Lc, rhs)
    cs = lm.ChainSolver(handle, M, pinned=True)
    cs.reduce(D, Lc, rhs)
    fixed_blocks = torch.zeros(M * 3, 25, dtype=torch.uint8, device="cuda")
    fixed_blocks[:N] = s["fixed"]
    ref = lm.pack_interface(D, Lc, rhs, fixed_blocks.view(M, 75)).cpu().numpy()
    sol.ctl[lm.CTL_LAM] = lam
    sol._enq(lm.PH_REDUCE)
    torch.cuda.synchronize()
    got = sol.payload.cpu().numpy()
    assert int(sol.info.item()) == 0
    n2 = 75 * 75
    for k, name in enumerate(("D_first", "D_last", "Lc_first", "Lc_last")):
        a, b = got[k * n2:(k + 1) * n2], ref[k * n2:(k + 1) * n2]
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1e-300), name
    assert np.abs(got[4 * n2:4 * n2 + 150] - ref[4 * n2:4 * n2 + 150]).max() <= 1e-9 * np.abs(ref[4 * n2:4 * n2 + 150]).max()
    assert np.array_equal(got[4 * n2 + 150:], ref[4 * n2 + 150:])
    if frame0 > 0:
        assert np.abs(ref[2 * n2:3 * n2]).max() > 0      # the coupling to the previous rank's last block is there


def test_fte_solve_graph_equals_eager(handle, dummy_cams):
    """The CUDA-graph replay of an attempt and the eager enqueue are the same kernels on the same data: identical bits."""
    import synth
    from acinoset_b200 import lm
    from oracle import fisheye, skeleton

    N = 120
    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=9, cams=dummy_cams)
    xa, ia = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"], use_graph=True).solve(p["x0"], max_iter=30)
    xb, ib = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"], use_graph=False).solve(p["x0"], max_iter=30)
    assert ia["graph"] and not ib["graph"]
    assert ia["n_solve"] == ib["n_solve"] and ia["iters"] == ib["iters"]
    assert ia["F"] == ib["F"] and np.array_equal(xa, xb)
    assert ia["converged"] and ia["history"][-1] == ia["F"]
    assert all(b <= a for a, b in zip(ia["history"], ia["history"][1:]))


@pytest.mark.parametrize("N", [48, 200])
def test_fte_solve_matches_cpu_restatement(handle, dummy_cams, N):
    """Solve parity: final objective within 1e-4 relative and marker positions within
    1e-3 m of the fp64 CPU restatement running the same algorithm."""
    import synth
    from acinoset_b200 import lm
    from oracle import fisheye, lm as olm, skeleton

    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=5, cams=dummy_cams)
    x_ref, info_ref = olm.solve(p, p["x0"], max_iter=40)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"])
    x, info = sol.solve(p["x0"], max_iter=40)
    assert info["bcr_info"] == 0
    assert abs(info["F"] - info_ref["F"]) < 1e-4 * abs(info_ref["F"])
    P, Pr = skeleton.cheetah_fk_active(x), skeleton.cheetah_fk_active(x_ref)
    assert np.abs(P - Pr).max() < 1e-3
    # and it actually solved the problem: close to the ground truth
    Pt = skeleton.cheetah_fk_active(p["x_true"])
    assert np.sqrt(((P - Pt) ** 2).sum(-1).mean()) < 0.01
    lo, hi = skeleton.active_bounds()
    assert (x >= lo - 1e-12).all() and (x <= hi + 1e-12).all()


@pytest.mark.parametrize("N,seed", [(48, 5), (100, 6)])
def test_fte_solve_reaches_independent_optimum(handle, dummy_cams, N, seed):
    """Solve pin against an INDEPENDENT optimiser (tests/golden/solves.npz: SciPy's More-Sorensen trust-region Newton on
    the oracle's fp64 objective from the same start, tests/golden/make_golden_solves.py): the GPU's end point, judged by
    the same fp64 objective, is at least as low (1e-6 relative) and its markers agree to 1e-3 m."""
    import synth
    from conftest import golden
    from acinoset_b200 import lm
    from oracle import fisheye, fte as ofte, skeleton

    g = golden("solves.npz")
    assert int(g[f"fte{N}_seed"]) == seed
    p = synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=seed, cams=dummy_cams)
    sol = lm.FTESolver(handle, p["meas"], p["w"], p["Ts"])
    x, info = sol.solve(p["x0"], max_iter=60)
    assert info["bcr_info"] == 0 and info["converged"]
    K, D, R, t, _ = dummy_cams
    meas = p["meas"].astype(np.float32).astype(np.float64)
    w = p["w"].astype(np.float32).astype(np.float64)
    q = ofte.model_weights_active()

    def F64(xx):
        c, _, _ = ofte.fte_eval(xx, meas, w, K, D, R, t)
        return float(c.sum()) + ofte.smooth_cost(xx, p["Ts"], q)

    F_scipy = float(g[f"fte{N}_F"])
    assert abs(F64(g[f"fte{N}_x"]) - F_scipy) < 1e-9 * F_scipy          # the fixture is what the script computed
    assert F64(x) <= F_scipy * (1 + 1e-6), (F64(x), F_scipy)
    P, Ps = skeleton.cheetah_fk_active(x), skeleton.cheetah_fk_active(g[f"fte{N}_x"])
    assert np.abs(P - Ps).max() < 1e-3, np.abs(P - Ps).max()


def test_solver_refuses_a_replaced_scene(dummy_cams):
    """A handle holds ONE camera table; a solver (whose captured graph carries the scene by value) must not silently run
    against another one that a TRI / SBA call installed in between."""
    import acinoset_b200 as ab
    import synth
    from acinoset_b200 import lm
    from oracle import fisheye, skeleton

    K, D, R, t, _ = dummy_cams
    h = ab.Handle(0)
    h.set_cameras(K, D, R, t)
    p = synth.make_fte_problem(12, skeleton.cheetah_fk_active, fisheye.project, seed=1, cams=dummy_cams)
    sol = lm.FTESolver(h, p["meas"], p["w"], p["Ts"])
    sol.solve(p["x0"], max_iter=2)
    h.set_cameras(K, D, R, t)                      # the same table again is not a change
    sol.solve(p["x0"], max_iter=2)
    h.set_cameras(K, D, R, t + 0.01)
    with pytest.raises(ab.AcinoError, match="camera table"):
        sol.solve(p["x0"], max_iter=2)
    sol.close()
    h.close()
