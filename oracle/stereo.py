"""Oracle (test infrastructure): pairwise extrinsic calibration of two fisheye cameras, fp64 NumPy / SciPy.

Follows calibrate_pair_extrinsics_fisheye (/root/reference/src/calib/calib.py:125-134), which hands the problem to
cv2.fisheye.stereoCalibrate(flags=CALIB_FIX_INTRINSIC, criteria=(MAX_ITER + EPS, 100, 1e-5)).  OpenCV is an
un-vendored dependency (conda_envs/acinoset.yml:13, unpinned; 4.13.0 in this image); its published objective, restated:
    minimise  sum_v sum_m |proj1(R_v X_m + t_v) - u1_vm|^2 + |proj2(R (R_v X_m + t_v) + T) - u2_vm|^2
over the relative pose (R, T) and one board pose (R_v, t_v) per view, rms = sqrt(objective / (2 V M)).
`solve` minimises it with scipy.optimize.least_squares over Rodrigues vectors - an independent route from the kernels'
multiplicative-update Levenberg-Marquardt (csrc/stereo_body.cuh).  Pinned by tests/golden/stereo.npz: cv2's own
(rms, R, T) on the reference's shipped checkerboard points, the RMS values its notebook prints
(calib_with_gui.ipynb:665,673) and the scene files those runs produced.
"""
import numpy as np

from . import fisheye


def residuals(rel_rvec, rel_t, rvecs, tvecs, obj, img1, img2, K1, D1, K2, D2):
    """-> (V, M, 4): (u1, v1, u2, v2) reprojection errors."""
    Rr = fisheye.rodrigues(rel_rvec)
    out = np.empty(img1.shape[:2] + (4,))
    for v in range(img1.shape[0]):
        X1 = obj @ fisheye.rodrigues(rvecs[v]).T + tvecs[v]
        out[v, :, :2] = fisheye.project(X1, K1, D1, np.eye(3), np.zeros(3)) - img1[v]
        out[v, :, 2:] = fisheye.project(X1, K2, D2, Rr, rel_t) - img2[v]
    return out


def rms(R, T, poses, obj, img1, img2, K1, D1, K2, D2):
    """poses (V, 12) = board poses in camera 1 (R row-major, t)."""
    rv = np.array([fisheye.rodrigues_inv(p[:9].reshape(3, 3)) for p in poses])
    e = residuals(fisheye.rodrigues_inv(R), np.asarray(T).reshape(3), rv, poses[:, 9:], obj, img1, img2, K1, D1, K2, D2)
    return float(np.sqrt((e ** 2).sum() / (2 * img1.shape[0] * img1.shape[1])))


def solve(R0, T0, poses0, obj, img1, img2, K1, D1, K2, D2):
    """scipy least_squares from an initial point -> (rms, R, T)."""
    from scipy.optimize import least_squares

    V, M = img1.shape[:2]
    x0 = np.concatenate([fisheye.rodrigues_inv(R0), np.asarray(T0).reshape(3)] +
                        [np.concatenate([fisheye.rodrigues_inv(p[:9].reshape(3, 3)), p[9:]]) for p in poses0])

    def fun(x):
        pv = x[6:].reshape(V, 6)
        return residuals(x[:3], x[3:6], pv[:, :3], pv[:, 3:], obj, img1, img2, K1, D1, K2, D2).ravel()

    res = least_squares(fun, x0, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    return float(np.sqrt((res.fun ** 2).sum() / (2 * V * M))), fisheye.rodrigues(res.x[:3]), res.x[3:6].reshape(3, 1)
