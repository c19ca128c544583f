"""Oracle (test infrastructure): two-view DLT triangulation and the pairwise TRI driver,
fp64 NumPy.

Follows
  * triangulate_points_fisheye            /root/reference/src/calib/calib.py:121-130
      cv2.fisheye.undistortPoints x2 -> P = [R|t] -> cv2.triangulatePoints -> dehomogenise
  * get_pairwise_3d_points_from_df        calib.py:394-423
      adjacent camera pairs (i,i+1) only, inner merge on (frame, marker), unweighted mean
      of the <= C-1 pairwise estimates.
cv2.triangulatePoints (OpenCV, un-vendored): per point the 4x4 homogeneous system
    A = [x1 P1[2]-P1[0]; y1 P1[2]-P1[1]; x2 P2[2]-P2[0]; y2 P2[2]-P2[1]]
(no row normalisation), X = right singular vector of the smallest singular value.
"""
import numpy as np

from . import fisheye


def dlt_pair(x1, x2, P1, P2):
    """x1,x2 (n,2) normalised image coordinates; P (3,4).  -> (n,3)."""
    x1 = np.asarray(x1, dtype=np.float64).reshape(-1, 2)
    x2 = np.asarray(x2, dtype=np.float64).reshape(-1, 2)
    out = np.empty((x1.shape[0], 3))
    for i in range(x1.shape[0]):
        A = np.stack([
            x1[i, 0] * P1[2] - P1[0],
            x1[i, 1] * P1[2] - P1[1],
            x2[i, 0] * P2[2] - P2[0],
            x2[i, 1] * P2[2] - P2[1],
        ])
        _, _, Vt = np.linalg.svd(A)
        X = Vt[-1]
        out[i] = X[:3] / X[3]
    return out


def triangulate_points_fisheye(img_pts_1, img_pts_2, k1, d1, r1, t1, k2, d2, r2, t2):
    p1 = fisheye.undistort(np.asarray(img_pts_1, dtype=np.float64).reshape(-1, 2), k1, d1)
    p2 = fisheye.undistort(np.asarray(img_pts_2, dtype=np.float64).reshape(-1, 2), k2, d2)
    P1 = np.hstack([np.asarray(r1, dtype=np.float64).reshape(3, 3), np.asarray(t1, dtype=np.float64).reshape(3, 1)])
    P2 = np.hstack([np.asarray(r2, dtype=np.float64).reshape(3, 3), np.asarray(t2, dtype=np.float64).reshape(3, 1)])
    return dlt_pair(p1, p2, P1, P2)


def pairwise_mean_dense(uv, valid, K, D, R, t):
    """Dense-tensor form of get_pairwise_3d_points_from_df.

    uv (N,C,L,2), valid (N,C,L) bool (the rows that survive the caller's likelihood
    filter).  Returns (pos (N,L,3) with NaN where no adjacent pair saw the point,
    count (N,L) number of pairs averaged).  Pair order 0-1, 1-2, ...; the mean is the
    sum in that order divided by the count.
    """
    uv = np.asarray(uv, dtype=np.float64)
    N, C, L, _ = uv.shape
    acc = np.zeros((N, L, 3))
    cnt = np.zeros((N, L), dtype=np.int64)
    for c in range(C - 1):
        both = valid[:, c] & valid[:, c + 1]
        idx = np.nonzero(both)
        if idx[0].size == 0:
            continue
        X = triangulate_points_fisheye(uv[:, c][idx], uv[:, c + 1][idx], K[c], D[c], R[c], t[c],
                                       K[c + 1], D[c + 1], R[c + 1], t[c + 1])
        acc[idx] += X
        cnt[idx] += 1
    pos = np.full((N, L, 3), np.nan)
    m = cnt > 0
    pos[m] = acc[m] / cnt[m][:, None]
    return pos, cnt
