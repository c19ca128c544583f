"""Oracle (test infrastructure): the standard (pinhole) camera model twins, fp64 NumPy.

Follows /root/reference/src/calib/calib.py
  * project_points                     :64-66   cv2.projectPoints(obj_pts, r, t, k, d)
  * create_undistort_point_function    :25-30   cv2.undistortPoints(pts, k, d, P=k)
  * triangulate_points                 :52-61   cv2.undistortPoints x2 -> [R|t] -> cv2.triangulatePoints
OpenCV is an un-vendored dependency (conda_envs/acinoset.yml:13, unpinned; 4.13.0 in this image).  Its
published model, restated: x' = X/Z, y' = Y/Z (Z == 0 -> divide by 1), r2 = x'^2 + y'^2,
    x'' = x' (1 + k1 r2 + k2 r4 + k3 r6)/(1 + k4 r2 + k5 r4 + k6 r6) + 2 p1 x'y' + p2 (r2 + 2 x'^2) + s1 r2 + s2 r4
    y'' = y' (...)/(...)                                          + p1 (r2 + 2 y'^2) + 2 p2 x'y' + s3 r2 + s4 r4
with d = [k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4].  cv2.undistortPoints inverts it with exactly 5 fixed-point
iterations (default TermCriteria(MAX_ITER, 5, 0.01): no convergence test).  Pinned to cv2 outputs in
tests/golden/pinhole.npz (generated through the reference's own functions).
"""
import numpy as np

from . import fisheye, triangulate


def _coeffs(d):
    d = np.zeros(0) if d is None else np.asarray(d, dtype=np.float64).ravel()
    if d.size > 12 and np.any(d[12:] != 0):
        raise ValueError("tilted-sensor terms are not restated")
    out = np.zeros(12)
    out[:min(d.size, 12)] = d[:12]
    return out


def project_points(obj_pts, k, d, r, t):
    X = np.asarray(obj_pts, dtype=np.float64).reshape(-1, 3)
    r = np.asarray(r, dtype=np.float64)
    Rm = fisheye.rodrigues(fisheye.rodrigues_inv(r.reshape(3, 3))) if r.size == 9 else fisheye.rodrigues(r.reshape(3))
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = _coeffs(d)
    Xc = X @ Rm.T + np.asarray(t, dtype=np.float64).reshape(1, 3)
    z = np.where(Xc[:, 2] != 0, 1.0 / np.where(Xc[:, 2] != 0, Xc[:, 2], 1.0), 1.0)
    x, y = Xc[:, 0] * z, Xc[:, 1] * z
    r2 = x * x + y * y
    r4, r6 = r2 * r2, r2 * r2 * r2
    cd = (1 + k1 * r2 + k2 * r4 + k3 * r6) / (1 + k4 * r2 + k5 * r4 + k6 * r6)
    xd = x * cd + 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + s1 * r2 + s2 * r4
    yd = y * cd + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + s3 * r2 + s4 * r4
    k = np.asarray(k, dtype=np.float64)
    return np.stack([k[0, 0] * xd + k[0, 2], k[1, 1] * yd + k[1, 2]], axis=1)


def undistort_points(pts, k, d, to_pixels=False, iters=5):
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    k = np.asarray(k, dtype=np.float64)
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = _coeffs(d)
    out = np.empty_like(pts)
    for i, (u, v) in enumerate(pts):
        x = (u - k[0, 2]) / k[0, 0]
        y = (v - k[1, 2]) / k[1, 1]
        x0, y0 = x, y
        for _ in range(iters):
            r2 = x * x + y * y
            icd = (1 + ((k6 * r2 + k5) * r2 + k4) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
            if icd < 0:
                x, y = x0, y0
                break
            dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + s1 * r2 + s2 * r2 * r2
            dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + s3 * r2 + s4 * r2 * r2
            x = (x0 - dx) * icd
            y = (y0 - dy) * icd
        out[i] = (x * k[0, 0] + k[0, 2], y * k[1, 1] + k[1, 2]) if to_pixels else (x, y)
    return out


def triangulate_points(img_pts_1, img_pts_2, k1, d1, r1, t1, k2, d2, r2, t2):
    p1 = undistort_points(img_pts_1, k1, d1)
    p2 = undistort_points(img_pts_2, k2, d2)
    P1 = np.hstack([np.asarray(r1, dtype=np.float64).reshape(3, 3), np.asarray(t1, dtype=np.float64).reshape(3, 1)])
    P2 = np.hstack([np.asarray(r2, dtype=np.float64).reshape(3, 3), np.asarray(t2, dtype=np.float64).reshape(3, 1)])
    return triangulate.dlt_pair(p1, p2, P1, P2)
