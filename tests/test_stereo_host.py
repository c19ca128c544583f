"""CPU: pairwise extrinsic calibration.  The kernel bodies (acinoset_b200/csrc/stereo_body.cuh) compiled for
the host by tests/host_harness/stereo_host.cpp, driven by the product's own accept / reject loop
(acinoset_b200/stereo.py), against cv2.fisheye.stereoCalibrate run the way the reference runs it on its shipped
checkerboard points (tests/golden/stereo.npz), the RMS values the reference's notebook prints, the scene files it
shipped, and the SciPy oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class HostBackend:
    """Same interface as acinoset_b200.stereo._GpuBackend on the host-compiled kernel bodies (test infrastructure)."""

    def __init__(self, lib):
        self.lib = lib

    def set(self, obj, img1, img2, K1, D1, K2, D2, pinhole=False):
        self.V, self.M = img1.shape[0], img1.shape[1]
        self.lib.stereo_host_set(self.V, self.M, _p(obj), _p(img1), _p(img2), _p(K1), _p(D1), min(D1.size, 12), _p(K2), _p(D2),
                                 min(D2.size, 12), 1 if pinhole else 0)

    def init(self):
        poses, cost = np.empty((self.V, 2, 12)), np.empty((self.V, 2))
        self.lib.stereo_host_init(_p(poses), _p(cost))
        return poses, cost

    def step(self, rel, poses, lam):
        rel_t, poses_t = np.empty(12), np.empty((self.V, 12))
        c, i = ctypes.c_double(), ctypes.c_int32()
        self.lib.stereo_host_step(_p(rel), _p(poses), ctypes.c_double(lam), _p(rel_t), _p(poses_t), ctypes.byref(c), ctypes.byref(i))
        return rel_t, poses_t, c.value, i.value


@pytest.fixture(scope="module")
def host_backend(tmp_path_factory):
    out = tmp_path_factory.mktemp("stereo_host") / "stereo_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                           os.path.join(ROOT, "tests", "host_harness", "stereo_host.cpp")])
    return HostBackend(ctypes.CDLL(str(out)))


@pytest.fixture
def backend(host_backend, monkeypatch):
    """The product has no backend switch: the tests swap the module's GPU backend class for the host harness."""
    from acinoset_b200 import stereo

    monkeypatch.setattr(stereo, "_GpuBackend", lambda device=0: host_backend)
    return host_backend


def check_against_reference(solve, tag, notebook_rms):
    """Shared with tests/test_stereo_gpu.py: `solve(obj, img1, img2, K1, D1, K2, D2) -> (rms, R, T, info)`."""
    g = golden("stereo.npz")
    a = [g[f"{tag}_{k}"] for k in ("obj", "img1", "img2", "K1", "D1", "K2", "D2")]
    rms, R, T, info = solve(*a)
    # cv2.fisheye.stereoCalibrate (stops at a relative change of 1e-5): same minimum, ours converged at least as far
    assert abs(rms - float(g[f"{tag}_rms"])) < 2e-5 and rms <= float(g[f"{tag}_rms"]) + 1e-9
    assert abs(rms - notebook_rms) < 2e-5                       # calib_with_gui.ipynb:665,673 (cv2 stopped a little earlier)
    assert np.abs(R - g[f"{tag}_R"]).max() < 1e-5 and np.abs(T - g[f"{tag}_T"]).max() < 1e-5
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    # chained with the first camera of the reference's shipped scene file it gives the second one (calib.py:186-187)
    Ra, ta = g[f"{tag}_scene_R"], g[f"{tag}_scene_t"]
    assert np.abs(R @ Ra[0] - Ra[1]).max() < 1e-5
    assert np.abs((R @ ta[0].reshape(3, 1) + T).ravel() - ta[1].ravel()).max() < 1e-5
    return rms, R, T, info, a


@pytest.mark.parametrize("tag,notebook_rms", [("rot12", 0.32182), ("sta34", 0.36876)])
def test_pair_calibration_matches_cv2_and_the_shipped_scenes(backend, tag, notebook_rms):
    from acinoset_b200 import stereo
    from oracle import stereo as ostereo

    def solve(*a):
        return stereo.solve_pair(*a, return_info=True)

    rms, R, T, info, (obj, img1, img2, K1, D1, K2, D2) = check_against_reference(solve, tag, notebook_rms)
    obj, img1, img2 = obj.astype(np.float64), img1.reshape(16, -1, 2).astype(np.float64), img2.reshape(16, -1, 2).astype(np.float64)
    # the oracle's restatement of the objective agrees with the kernels' cost at the solution ...
    assert abs(ostereo.rms(R, T, info["poses"], obj, img1, img2, K1, D1, K2, D2) - rms) < 1e-9
    # ... and an independent SciPy minimisation from there does not find anything better
    rms_o, R_o, T_o = ostereo.solve(R, T, info["poses"], obj, img1, img2, K1, D1, K2, D2)
    assert rms_o <= rms + 1e-12 and rms - rms_o < 1e-7
    assert np.abs(R_o - R).max() < 1e-5 and np.abs(T_o - T).max() < 1e-5
    # the per-view initialisation (homography + single-camera refinement) is already sub-pixel
    assert np.all(info["init_cost"] >= 0) and np.sqrt(info["init_cost"].sum() / (2 * 16 * 54 * 2)) < 0.3


def test_synthetic_pair_recovers_ground_truth(backend):
    """Noise-free synthetic boards: exact recovery; 0.2 px noise: rms ~ noise level."""
    from acinoset_b200 import stereo, utils
    from acinoset_b200.rotations import rodrigues_to_mat
    from oracle import fisheye

    g = golden("stereo.npz")
    K1, D1, K2, D2 = g["rot12_K1"], g["rot12_D1"], g["rot12_K2"], g["rot12_D2"]
    rng = np.random.default_rng(5)
    obj = utils.create_board_object_pts((9, 6), 0.05).astype(np.float64)
    R_true = rodrigues_to_mat(np.array([0.05, -0.4, 0.02]))
    T_true = np.array([-0.6, 0.03, 0.1])
    V = 12
    img1, img2 = np.empty((V, 54, 2)), np.empty((V, 54, 2))
    for v in range(V):
        Rv = rodrigues_to_mat(rng.normal(0, 0.3, 3))
        tv = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(0.8, 1.6)])
        X1 = (obj - obj.mean(0)) @ Rv.T + tv
        img1[v] = fisheye.project(X1, K1, D1, np.eye(3), np.zeros(3))
        img2[v] = fisheye.project(X1, K2, D2, R_true, T_true)
    obj_c = obj - obj.mean(0)
    rms, R, T = stereo.solve_pair(obj_c, img1, img2, K1, D1, K2, D2)
    assert rms < 1e-6 and np.abs(R - R_true).max() < 1e-8 and np.abs(T.ravel() - T_true).max() < 1e-8
    rms_n, R_n, T_n = stereo.solve_pair(obj_c, img1 + rng.normal(0, 0.2, img1.shape), img2 + rng.normal(0, 0.2, img2.shape),
                                        K1, D1, K2, D2)
    assert 0.1 < rms_n < 0.3 and np.abs(R_n - R_true).max() < 2e-3 and np.abs(T_n.ravel() - T_true).max() < 5e-3


def test_pairwise_chain_with_injected_calib_func():
    """calibrate_pairwise_extrinsics (calib.py:141-194): view matching by file name and pose chaining (host logic only)."""
    from acinoset_b200 import stereo
    from acinoset_b200.rotations import rodrigues_to_mat

    rel = [(rodrigues_to_mat(np.array([0.0, 0.3, 0.0])), np.array([[-1.0], [0.0], [0.1]])),
           (rodrigues_to_mat(np.array([0.1, 0.2, -0.1])), np.array([[-0.8], [0.1], [0.0]]))]
    seen = []

    def fake(obj_pts, p1, p2, k1, d1, k2, d2, res):
        seen.append((p1.shape, p2.shape, float(p1[0, 0, 0, 0]), float(p2[0, 0, 0, 0])))
        return (0.25,) + rel[len(seen) - 1]

    pts = [np.full((3, 9, 6, 2), 10.0), np.full((4, 9, 6, 2), 20.0), np.full((2, 9, 6, 2), 30.0)]
    for c in range(3):
        pts[c] += np.arange(len(pts[c]))[:, None, None, None]
    names = [["a.jpg", "b.jpg", "c.jpg"], ["c.jpg", "x.jpg", "a.jpg", "y.jpg"], ["y.jpg", "c.jpg"]]
    r_arr, t_arr = stereo.calibrate_pairwise_extrinsics(fake, pts, names, [np.eye(3)] * 3, [np.zeros(4)] * 3, (10, 10), (9, 6), 0.03)
    assert seen[0][:2] == ((2, 9, 6, 2), (2, 9, 6, 2)) and seen[0][2:] == (10.0, 22.0)     # a.jpg: view 0 of cam 1, view 2 of cam 2
    assert seen[1][:2] == ((2, 9, 6, 2), (2, 9, 6, 2)) and seen[1][2:] == (20.0, 31.0)     # c.jpg: view 0 of cam 2, view 1 of cam 3
    R1 = np.array([[1.0, 0, 0], [0, 0, -1], [0, 1, 0]])
    assert np.array_equal(r_arr[0], R1) and np.array_equal(t_arr[0], np.zeros((3, 1)))
    assert np.allclose(r_arr[1], rel[0][0] @ R1) and np.allclose(t_arr[1], rel[0][1])
    assert np.allclose(r_arr[2], rel[1][0] @ rel[0][0] @ R1) and np.allclose(t_arr[2], rel[1][0] @ rel[0][1] + rel[1][1])
    with pytest.raises(AssertionError):
        stereo.calibrate_pairwise_extrinsics(fake, pts[:2], [["a"], ["b"]], [np.eye(3)] * 2, [np.zeros(4)] * 2, (10, 10), (9, 6), 0.03)


def check_pinhole_against_reference(pair_func):
    """Shared with the GPU test: the reference's calibrate_pair_extrinsics (cv2.stereoCalibrate, run unmodified) on
    noise-free synthetic boards seen through the 8-coefficient rational model of calib.py:18.
    pair_func(obj, img1, img2, k1, d1, k2, d2, resolution, rational_model=...) -> (rms, r, t)."""
    g = golden("stereo.npz")
    a = (g["pin_obj"], g["pin_img1"].reshape(-1, 9, 6, 2), g["pin_img2"].reshape(-1, 9, 6, 2), g["pin_K1"], g["pin_D1"], g["pin_K2"],
         g["pin_D2"], (1920, 1080))
    # default = the reference's behaviour: OpenCV ignores k4..k6 without CALIB_RATIONAL_MODEL -> a 0.49 px model error
    rms, R, T = pair_func(*a, rational_model=False)
    assert abs(rms - float(g["pin_rms"])) < 1e-6
    assert np.abs(R - g["pin_R"]).max() < 1e-6 and np.abs(T - g["pin_T"]).max() < 1e-6 and T.shape == (3, 1)
    # every coefficient used: the noise-free ground truth comes back (pixels were rounded to float32)
    rms_f, R_f, T_f = pair_func(*a, rational_model=True)
    assert rms_f < 1e-4 and np.abs(R_f - g["pin_R_true"]).max() < 1e-7 and np.abs(T_f.ravel() - g["pin_T_true"]).max() < 1e-6


def test_pinhole_pair_calibration_matches_reference(backend):
    from acinoset_b200 import stereo

    check_pinhole_against_reference(lambda *a, **k: stereo.calibrate_pair_extrinsics(*a, **k))
