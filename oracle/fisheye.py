"""Oracle (test infrastructure): Kannala-Brandt fisheye camera model in fp64 NumPy.

Follows
  * ``pt3d_to_2d``                  /root/reference/src/all_optimizations.py:193-209
                                     (= src/build.py:457-473)
  * ``project_points_fisheye``      /root/reference/src/calib/calib.py:132-136
                                     (cv2.Rodrigues + cv2.fisheye.projectPoints)
  * ``cv2.fisheye.undistortPoints`` as called at calib.py:124-125 (no R / P given,
                                     default criteria) - OpenCV is an un-vendored,
                                     un-pinned dependency (conda_envs/acinoset.yml:13);
                                     restated from its published algorithm and pinned
                                     against cv2 4.13.0 outputs in tests/golden/.
  * ``cv2.Rodrigues``               calib.py:134,349,373
"""
import numpy as np

SQRT_EPS_R2 = 1e-12  # the "+1e-12" under the square root, all_optimizations.py:201


def _split_cam(K, D, R, t):
    K = np.asarray(K, dtype=np.float64)
    D = np.asarray(D, dtype=np.float64).reshape(-1)
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(3)
    return K, D, R, t


def theta_d(th, D):
    """th * (1 + D0 th^2 + D1 th^4 + D2 th^6 + D3 th^8)  (all_optimizations.py:204)."""
    th2 = th * th
    return th * (1.0 + th2 * (D[0] + th2 * (D[1] + th2 * (D[2] + th2 * D[3]))))


def dtheta_d(th, D):
    th2 = th * th
    return 1.0 + th2 * (3 * D[0] + th2 * (5 * D[1] + th2 * (7 * D[2] + th2 * 9 * D[3])))


def project(X, K, D, R, t):
    """pt3d_to_2d for an array of world points X (...,3) -> (...,2).

    No skew, no behind-camera guard - exactly like the reference.
    """
    K, D, R, t = _split_cam(K, D, R, t)
    X = np.asarray(X, dtype=np.float64)
    Xc = X @ R.T + t
    a = Xc[..., 0] / Xc[..., 2]
    b = Xc[..., 1] / Xc[..., 2]
    r = np.sqrt(a * a + b * b + SQRT_EPS_R2)
    th = np.arctan(r)
    s = theta_d(th, D) / r
    u = K[0, 0] * a * s + K[0, 2]
    v = K[1, 1] * b * s + K[1, 2]
    return np.stack([u, v], axis=-1)


def project_jac(X, K, D, R, t):
    """Projection and its Jacobian.

    Returns (uv (...,2), J_world (...,2,3) = d(u,v)/dX_world, J_cam (...,2,3) =
    d(u,v)/dX_cam).  Closed form of SURVEY.md appendix B2.
    """
    K, D, R, t = _split_cam(K, D, R, t)
    X = np.asarray(X, dtype=np.float64)
    Xc = X @ R.T + t
    x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    iz = 1.0 / z
    a = x * iz
    b = y * iz
    r2 = a * a + b * b + SQRT_EPS_R2
    r = np.sqrt(r2)
    th = np.arctan(r)
    td = theta_d(th, D)
    dtd = dtheta_d(th, D)
    s = td / r
    # ds/dr = (dtd * dth/dr * r - td) / r^2,  dth/dr = 1/(1+r^2)
    dsdr = (dtd * r / (1.0 + r * r) - td) / r2
    q = dsdr / r
    fx, fy = K[0, 0], K[1, 1]
    u = fx * a * s + K[0, 2]
    v = fy * b * s + K[1, 2]
    # d(a s, b s)/d(a, b)
    m00 = s + a * a * q
    m01 = a * b * q
    m11 = s + b * b * q
    # d(a,b)/dXc
    Jc = np.zeros(X.shape[:-1] + (2, 3))
    Jc[..., 0, 0] = fx * m00 * iz
    Jc[..., 0, 1] = fx * m01 * iz
    Jc[..., 0, 2] = -fx * (m00 * a + m01 * b) * iz
    Jc[..., 1, 0] = fy * m01 * iz
    Jc[..., 1, 1] = fy * m11 * iz
    Jc[..., 1, 2] = -fy * (m01 * a + m11 * b) * iz
    Jw = Jc @ R
    return np.stack([u, v], axis=-1), Jw, Jc


def rodrigues(rvec):
    """Rotation vector -> matrix (cv2.Rodrigues forward direction)."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = np.linalg.norm(rvec)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def rodrigues_inv(Rm):
    """Rotation matrix -> vector (what project_points_fisheye does at calib.py:134)."""
    Rm = np.asarray(Rm, dtype=np.float64).reshape(3, 3)
    # project to SO(3) like OpenCV does (SVD) before extracting the axis
    U, _, Vt = np.linalg.svd(Rm)
    Rm = U @ Vt
    rx = Rm[2, 1] - Rm[1, 2]
    ry = Rm[0, 2] - Rm[2, 0]
    rz = Rm[1, 0] - Rm[0, 1]
    s = np.sqrt((rx * rx + ry * ry + rz * rz) * 0.25)
    c = np.clip((np.trace(Rm) - 1) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t0 = (Rm[0, 0] + 1) * 0.5
        t1 = (Rm[1, 1] + 1) * 0.5
        t2 = (Rm[2, 2] + 1) * 0.5
        vx = np.sqrt(max(t0, 0.0))
        vy = np.sqrt(max(t1, 0.0)) * (-1.0 if Rm[0, 1] < 0 else 1.0)
        vz = np.sqrt(max(t2, 0.0)) * (-1.0 if Rm[0, 2] < 0 else 1.0)
        if abs(vx) < abs(vy) and abs(vx) < abs(vz) and (Rm[1, 2] > 0) != (vy * vz > 0):
            vz = -vz
        v = np.array([vx, vy, vz])
        return v * (th / np.linalg.norm(v))
    vth = 1.0 / (2 * s) * th
    return np.array([rx, ry, rz]) * vth


def drodrigues(rvec, eps_small=1e-8):
    """dR/drvec as (3,3,3): out[i,j,k] = dR[i,j]/drvec[k]  (closed form).

    For theta -> 0 this tends to the skew generators.
    """
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th2 = r @ r
    out = np.zeros((3, 3, 3))
    gen = np.zeros((3, 3, 3))
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1
        gen[:, :, k] = np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
    if th2 < eps_small ** 2:
        return gen
    th = np.sqrt(th2)
    Rm = rodrigues(r)
    rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    I = np.eye(3)
    # Gallego & Yezzi 2015: dR/dr_k = (r_k [r]x + [r x (I-R) e_k]x) / |r|^2  R
    for k in range(3):
        e = I[:, k]
        w = np.cross(r, (I - Rm) @ e)
        wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        out[:, :, k] = ((r[k] * rx + wx) / th2) @ Rm
    return out


def undistort(pts, K, D, max_iter=10, eps=1e-8):
    """cv2.fisheye.undistortPoints(pts, K, D) -> normalised coords, default criteria.

    pw = ((u-cx)/fx, (v-cy)/fy); theta_d = clip(|pw|, -pi/2, pi/2); Newton on
    theta(1+k1 th^2+...) = theta_d from theta = theta_d, <= 10 iterations, stop when
    |fix| < 1e-8; scale = tan(theta)/theta_d.  Non-converged (or sign-flipped)
    solutions come back as (-1e6, -1e6).  
    """
    K = np.asarray(K, dtype=np.float64)
    D = np.asarray(D, dtype=np.float64).reshape(-1)
    pts = np.asarray(pts, dtype=np.float64)
    shp = pts.shape
    p = pts.reshape(-1, 2)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    pw = np.stack([(p[:, 0] - cx) / fx, (p[:, 1] - cy) / fy], axis=-1)
    out = np.empty_like(pw)
    for i in range(pw.shape[0]):
        thd = np.sqrt(pw[i, 0] ** 2 + pw[i, 1] ** 2)
        thd = min(max(-np.pi / 2, thd), np.pi / 2)
        converged = False
        th = thd
        scale = 0.0
        if thd > eps:
            for _ in range(max_iter):
                th2 = th * th
                th4 = th2 * th2
                th6 = th4 * th2
                th8 = th6 * th2
                k0t, k1t, k2t, k3t = D[0] * th2, D[1] * th4, D[2] * th6, D[3] * th8
                fix = (th * (1 + k0t + k1t + k2t + k3t) - thd) / (1 + 3 * k0t + 5 * k1t + 7 * k2t + 9 * k3t)
                th = th - fix
                if abs(fix) < eps:
                    converged = True
                    break
            scale = np.tan(th) / thd
        else:
            converged = True
        flipped = (thd < 0 and th > 0) or (thd > 0 and th < 0)
        if converged and not flipped:
            out[i] = pw[i] * scale
        else:
            out[i] = (-1000000.0, -1000000.0)
    return out.reshape(shp)
