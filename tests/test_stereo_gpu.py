"""GPU: pairwise extrinsic calibration through the C ABI (acino_stereo_*) vs cv2 / the reference's shipped artefacts."""
import json

import numpy as np
import pytest

from conftest import golden
from test_stereo_host import check_against_reference, check_pinhole_against_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,notebook_rms", [("rot12", 0.32182), ("sta34", 0.36876)])
def test_pair_calibration_matches_cv2_and_the_shipped_scenes(tag, notebook_rms):
    from acinoset_b200 import stereo

    def solve(*a):
        return stereo.solve_pair(*a, return_info=True)

    rms, R, T, info, a = check_against_reference(solve, tag, notebook_rms)
    # reference signature: (n_views, rows, cols, 2) float32 points, camera_resolution positional
    g = golden("stereo.npz")
    rms2, R2, T2 = stereo.calibrate_pair_extrinsics_fisheye(a[0], a[1], a[2], a[3], a[4].reshape(4, 1), a[5], a[6].reshape(4, 1),
                                                            tuple(g[f"{tag}_res"]))
    assert rms2 == rms and np.array_equal(R2, R) and T2.shape == (3, 1)


def test_calibrate_fisheye_extrinsics_pairwise_writes_the_scene(tmp_path):
    """app.calibrate_fisheye_extrinsics_pairwise (app.py:84-124) on files in the reference's formats."""
    from acinoset_b200 import stereo, utils

    g = golden("stereo.npz")
    res = tuple(int(v) for v in g["sta34_res"])
    cams, pts = [], []
    for i, (k, d, img) in enumerate([(g["sta34_K1"], g["sta34_D1"], g["sta34_img1"]), (g["sta34_K2"], g["sta34_D2"], g["sta34_img2"])]):
        cf, pf = str(tmp_path / f"camera_{i}.json"), str(tmp_path / f"points_{i}.json")
        utils.save_camera(cf, res, k, d)
        names = [f"img{j:05d}.jpg" for j in range(len(img))]
        if i == 1:                                   # different order + an extra view only camera 2 saw
            order = list(range(len(img)))[::-1]
            utils.save_points(pf, np.concatenate([img[order], img[:1] + 5.0]), [names[j] for j in order] + ["only2.jpg"], (9, 6), 0.031, res)
        else:
            utils.save_points(pf, img, names, (9, 6), 0.031, res)
        cams.append(cf)
        pts.append(pf)
    out = str(tmp_path / "2_cam_scene.json")
    r_arr, t_arr = stereo.calibrate_fisheye_extrinsics_pairwise(cams, pts, out)
    k_arr, d_arr, r_s, t_s, res_s = utils.load_scene(out)
    R1 = np.array([[1.0, 0, 0], [0, 0, -1], [0, 1, 0]])
    assert np.allclose(r_s[0], R1) and np.allclose(t_s[0], 0)
    assert np.abs(r_s[1] - g["sta34_R"] @ R1).max() < 1e-5 and np.abs(t_s[1].ravel() - g["sta34_T"].ravel()).max() < 1e-5
    assert tuple(res_s) == res and np.allclose(k_arr[1], g["sta34_K2"])


def test_pinhole_pair_calibration_matches_reference():
    """calibrate_pair_extrinsics (calib.py:41-49) by name, incl. the reference's dropped rational coefficients."""
    from acinoset_b200 import stereo

    check_pinhole_against_reference(stereo.calibrate_pair_extrinsics)
