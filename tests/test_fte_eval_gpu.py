"""Parity of the CUDA fte_eval / fk_project kernels with the fp64 oracle (through the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# fp64 -> fp32 tolerances
TOL_PX = 2e-3          # |du|, |dv| for theta < 60 deg
TOL_POS = 5e-6         # marker positions, m
TOL_REL_FRO = 1e-4     # relative Frobenius error of g and H per frame
TOL_COST = 2e-5        # relative, per frame (cost ~ 1e2..1e3)


@pytest.fixture(scope="module")
def handle(dummy_cams):
    import acinoset_b200 as ab

    K, D, R, t, _ = dummy_cams
    h = ab.Handle(0)
    h.set_cameras(K, D, R, t)
    yield h
    h.close()


def _problem(N, seed, cams):
    import synth
    from oracle import fisheye, skeleton

    return synth.make_fte_problem(N, skeleton.cheetah_fk_active, fisheye.project, seed=seed, cams=cams)


def test_fk_project_matches_oracle(handle, dummy_cams):
    from oracle import fte, skeleton

    p = _problem(200, 3, dummy_cams)
    K, D, R, t, _ = dummy_cams
    pos, uv = handle.fk_project(p["x0"])
    pos_ref = skeleton.cheetah_fk_active(p["x0"].astype(np.float32).astype(np.float64))
    assert np.abs(pos - pos_ref).max() < TOL_POS
    uv_ref = fte.reproject(p["x0"].astype(np.float32).astype(np.float64), K, D, R, t)
    ok = p["lik"] > 0  # in front of the camera, theta < 60 deg, inside the image (at x_true)
    assert np.abs(uv - uv_ref)[ok].max() < TOL_PX


@pytest.mark.parametrize("N,seed", [(1, 0), (16, 1), (17, 2), (333, 4), (1000, 5)])
def test_fte_eval_matches_oracle(handle, dummy_cams, N, seed):
    from oracle import fte

    p = _problem(N, seed, dummy_cams)
    K, D, R, t, _ = dummy_cams
    x32 = p["x0"].astype(np.float32)
    m32 = p["meas"].astype(np.float32)
    w32 = p["w"].astype(np.float32)
    cost, g, Hu = handle.fte_eval(x32, m32, w32)
    c_ref, g_ref, H_ref = fte.fte_eval(x32.astype(np.float64), m32.astype(np.float64), w32.astype(np.float64), K, D, R, t)
    Hu_ref = fte.pack_upper(H_ref)
    assert np.all(np.isfinite(cost)) and np.all(np.isfinite(g)) and np.all(np.isfinite(Hu))
    assert (np.abs(cost - c_ref) / np.maximum(1.0, np.abs(c_ref))).max() < TOL_COST
    gerr = np.linalg.norm(g - g_ref, axis=1) / np.maximum(1e-3, np.linalg.norm(g_ref, axis=1))
    assert gerr.max() < TOL_REL_FRO, gerr.max()
    H = fte.unpack_upper(Hu.astype(np.float64))
    herr = np.linalg.norm(H - H_ref, axis=(1, 2)) / np.linalg.norm(H_ref, axis=(1, 2))
    assert herr.max() < TOL_REL_FRO, herr.max()


def test_fte_eval_zero_weight_rows_ignore_measurements(handle, dummy_cams):
    p = _problem(40, 9, dummy_cams)
    x32 = p["x0"].astype(np.float32)
    m32 = p["meas"].astype(np.float32)
    w32 = p["w"].astype(np.float32)
    ref = handle.fte_eval(x32, m32, w32)
    m2 = m32.copy()
    m2[w32 == 0] = np.nan  # garbage where the weight is zero must not matter
    out = handle.fte_eval(x32, m2, w32)
    for a, b in zip(ref, out):
        assert np.array_equal(a, b)


def test_fte_eval_without_H_and_device_api(handle, dummy_cams):
    import torch

    p = _problem(100, 10, dummy_cams)
    x32 = p["x0"].astype(np.float32)
    m32 = p["meas"].astype(np.float32)
    w32 = p["w"].astype(np.float32)
    cost, g, Hu = handle.fte_eval(x32, m32, w32)
    cost2, g2, none = handle.fte_eval(x32, m32, w32, want_H=False)
    assert none is None
    assert np.array_equal(cost, cost2) and np.array_equal(g, g2)
    dev = torch.device("cuda:0")
    xd, md, wd = (torch.from_numpy(a).to(dev) for a in (x32, m32, w32))
    cd = torch.empty(100, device=dev)
    gd = torch.empty(100, 25, device=dev)
    Hd = torch.empty(100, 325, device=dev)
    handle.fte_eval_dev(xd, md, wd, cd, gd, Hd)
    torch.cuda.synchronize()
    assert np.array_equal(cd.cpu().numpy(), cost)
    assert np.array_equal(gd.cpu().numpy(), g)
    assert np.array_equal(Hd.cpu().numpy(), Hu)


def test_fte_eval_deterministic(handle, dummy_cams):
    p = _problem(257, 12, dummy_cams)
    a = handle.fte_eval(p["x0"], p["meas"], p["w"])
    b = handle.fte_eval(p["x0"], p["meas"], p["w"])
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
