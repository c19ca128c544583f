"""Host-side rotation-parameter conversions with cv2.Rodrigues semantics (both directions).

cv2.Rodrigues is what the reference uses to move between rotation matrices and the 3-vector
parameterisation of the SBA problem (/root/reference/src/calib/calib.py:134,349,373); a matrix input
is first projected onto SO(3) (SVD) exactly like OpenCV does.  Pinned against cv2 in
tests/test_oracle_golden.py (the oracle holds the same formulas) and tests/test_abi.py."""
import numpy as np


def rodrigues_to_vec(Rm):
    Rm = np.asarray(Rm, dtype=np.float64).reshape(3, 3)
    U, _, Vt = np.linalg.svd(Rm)
    Rm = U @ Vt
    r = np.array([Rm[2, 1] - Rm[1, 2], Rm[0, 2] - Rm[2, 0], Rm[1, 0] - Rm[0, 1]])
    s = np.sqrt((r @ r) * 0.25)
    c = np.clip((np.trace(Rm) - 1) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        v = np.sqrt(np.maximum((np.diag(Rm) + 1) * 0.5, 0.0))
        if Rm[0, 1] < 0:
            v[1] = -v[1]
        if Rm[0, 2] < 0:
            v[2] = -v[2]
        if abs(v[0]) < abs(v[1]) and abs(v[0]) < abs(v[2]) and (Rm[1, 2] > 0) != (v[1] * v[2] > 0):
            v[2] = -v[2]
        return v * (th / np.linalg.norm(v))
    return r * (th / (2 * s))


def rodrigues_to_mat(rvec):
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = np.linalg.norm(r)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


