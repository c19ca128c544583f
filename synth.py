"""Synthetic workload generator for tests and bench.py.

Pure data generation: the ground-truth trajectory, noise / outlier / likelihood patterns
and the synthetic SBA checkerboard scene.  Forward kinematics and camera projection are
injected as callables (the oracle in the parity tests, the CUDA path in bench.py) so this
module depends on neither.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DUMMY_SCENE = os.path.join(_HERE, "acinoset_b200", "data", "dummy_scene.json")

FPS = 120.0
N_ACTIVE = 25
# active-state slot order: x,y,z, phi0,phi1,phi3, theta0..13, psi0,psi1,psi3,psi4,psi5
_PI = np.pi
_BOUNDS = {  # active slot -> (lo, hi); reference all_optimizations.py:403-483
    3: (-_PI / 6, _PI / 6), 4: (-_PI / 6, _PI / 6), 5: (-_PI / 6, _PI / 6),
    6: (-_PI / 6, _PI / 6), 7: (-_PI / 6, _PI / 6), 8: (-_PI / 6, _PI / 6), 9: (-_PI / 6, _PI / 6),
    10: (-_PI / 1.5, _PI / 1.5), 11: (-_PI / 1.5, _PI / 1.5),
    12: (-_PI / 2, _PI / 2), 13: (-_PI, 0.0), 14: (-_PI / 2, _PI / 2), 15: (-_PI, 0.0),
    16: (-_PI / 2, _PI / 2), 17: (0.0, _PI), 18: (-_PI / 2, _PI / 2), 19: (0.0, _PI),
    21: (-_PI / 6, _PI / 6), 22: (-_PI / 6, _PI / 6), 23: (-_PI / 1.5, _PI / 1.5), 24: (-_PI / 1.5, _PI / 1.5),
}
PSI0_SLOT = 20


def load_dummy_scene(path=DUMMY_SCENE):
    """6-camera fisheye scene (configs/dummy_scene.json of the reference, verbatim)."""
    with open(path) as f:
        d = json.load(f)
    K = np.array([c["k"] for c in d["cameras"]], dtype=np.float64)
    D = np.array([c["d"] for c in d["cameras"]], dtype=np.float64).reshape(-1, 4)
    R = np.array([c["r"] for c in d["cameras"]], dtype=np.float64)
    t = np.array([c["t"] for c in d["cameras"]], dtype=np.float64).reshape(-1, 3)
    return K, D, R, t, tuple(d["camera_resolution"])


def make_trajectory(N, rng, fps=FPS, start=0):
    """Ground-truth active states (N,25): head on a 3 m circle centred (2.0,6.5,0.6),
    period 600 frames, psi0 tangent to the path, z += 0.05 sin(2 pi 3 t); every bounded
    angle = centre + 0.6 * half-range * sin(2 pi f t + phase), f ~ U(1,4) Hz."""
    n = np.arange(start, start + N, dtype=np.float64)
    tt = n / fps
    ang = 2 * np.pi * n / 600.0
    xa = np.zeros((N, N_ACTIVE))
    xa[:, 0] = 2.0 + 3.0 * np.cos(ang)
    xa[:, 1] = 6.5 + 3.0 * np.sin(ang)
    xa[:, 2] = 0.6 + 0.05 * np.sin(2 * np.pi * 3 * tt)
    # heading tangent to the circle (direction of travel); the cheetah's body extends
    # along -x_body so the head leads.  Unwrapped so the trajectory is smooth.
    xa[:, PSI0_SLOT] = ang + np.pi / 2
    for p, (lo, hi) in sorted(_BOUNDS.items()):
        f = rng.uniform(1.0, 4.0)
        ph = rng.uniform(0.0, 2 * np.pi)
        xa[:, p] = 0.5 * (lo + hi) + 0.6 * 0.5 * (hi - lo) * np.sin(2 * np.pi * f * tt + ph)
    return xa


def make_measurements(P, cams, project, rng, noise_px=2.0, outlier_frac=0.05, low_lik_frac=0.10,
                      max_theta_deg=60.0, uv_all=None):
    """P (N,L,3) world marker positions -> meas (N,C,L,2), likelihood (N,C,L).

    project(P, K, D, R, t) -> (N,L,2) (or pass uv_all (N,C,L,2) precomputed, e.g. by the
    fk_project CUDA kernel).  N(0, noise_px) noise; outlier_frac of (n,c,l)
    replaced by uniform-in-image outliers; likelihood ~ U(0.5,1) except low_lik_frac set
    to U(0,0.5); any point with theta > max_theta, behind the camera or outside the image
    gets likelihood 0.
    """
    K, D, R, t, res = cams
    N, L, _ = P.shape
    C = len(K)
    meas = np.zeros((N, C, L, 2))
    lik = rng.uniform(0.5, 1.0, (N, C, L))
    low = rng.uniform(0, 1, (N, C, L)) < low_lik_frac
    lik[low] = rng.uniform(0.0, 0.5, int(low.sum()))
    for c in range(C):
        if uv_all is not None:
            uv = np.asarray(uv_all[:, c], dtype=np.float64)
        else:
            uv = np.asarray(project(P, K[c], D[c], R[c], t[c]), dtype=np.float64)
        Xc = P @ R[c].T + t[c]
        theta = np.arctan2(np.hypot(Xc[..., 0], Xc[..., 1]), Xc[..., 2])
        ok = (theta < np.deg2rad(max_theta_deg)) & (Xc[..., 2] > 0)
        ok &= (uv[..., 0] >= 0) & (uv[..., 0] < res[0]) & (uv[..., 1] >= 0) & (uv[..., 1] < res[1])
        uv = uv + rng.normal(0.0, noise_px, uv.shape)
        out = rng.uniform(0, 1, (N, L)) < outlier_frac
        uv[out] = rng.uniform([0, 0], [res[0], res[1]], (int(out.sum()), 2))
        uv[~ok] = 0.0
        meas[:, c] = uv
        lik[:, c][~ok] = 0.0
    return meas, lik


def make_fte_problem(N, fk, project, seed=0, dlc_thresh=0.5, init_sigma=0.05, cams=None, start=0,
                     reproject=None):
    """Config 2/3/5 of BASELINE.json: returns dict(x_true, x0, meas, lik, w, cams, Ts).

    Either (fk, project) callables, or reproject(x) -> (P (N,L,3), uv (N,C,L,2))."""
    rng = np.random.default_rng(seed)
    cams = load_dummy_scene() if cams is None else cams
    x_true = make_trajectory(N, rng, start=start)
    if reproject is not None:
        P, uv_all = reproject(x_true)
        meas, lik = make_measurements(np.asarray(P, dtype=np.float64), cams, None, rng, uv_all=uv_all)
    else:
        P = np.asarray(fk(x_true), dtype=np.float64)
        meas, lik = make_measurements(P, cams, project, rng)
    w = np.where(lik > dlc_thresh, 1.0 / 5.0, 0.0)  # all_optimizations.py:243,302-308
    x0 = x_true + rng.normal(0.0, init_sigma, x_true.shape)
    return dict(x_true=x_true, x0=x0, meas=meas, lik=lik, w=w, cams=cams, Ts=1.0 / FPS,
                dlc_thresh=dlc_thresh)


def dense_to_long_df(meas, lik, markers):
    """Long-form DataFrame [frame,camera,marker,x,y,likelihood] (reference utils.py:105-120)."""
    import pandas as pd

    N, C, L, _ = meas.shape
    n, c, l = np.meshgrid(np.arange(N), np.arange(C), np.arange(L), indexing="ij")
    return pd.DataFrame({
        "frame": n.ravel(), "camera": c.ravel(),
        "marker": np.asarray(markers, dtype=object)[l.ravel()],
        "x": meas[..., 0].ravel(), "y": meas[..., 1].ravel(), "likelihood": lik.ravel(),
    })


def make_sba_problem(n_views, project, seed=0, board_shape=(9, 6), square=0.10, noise_px=0.2,
                     rot_sigma_deg=0.5, trans_sigma=0.02, cams=None, max_tries_factor=20):
    """Config 4: synthetic checkerboard views in the 6-camera dummy scene.

    Each view is placed 2-6 m in front of the midpoint of an adjacent camera pair with
    <= 40 deg tilt and kept if >= 2 cameras see all corners at theta < 60 deg inside the
    image.  Returns dict with ground truth, observations in the reference's flat layout
    (points_2d (n_obs,2) f32, point_3d_indices, camera_indices) and perturbed extrinsics.
    """
    rng = np.random.default_rng(seed)
    K, D, R, t, res = load_dummy_scene() if cams is None else cams
    C = len(K)
    n_pts = board_shape[0] * board_shape[1]
    gx, gy = np.meshgrid(np.arange(board_shape[0]), np.arange(board_shape[1]), indexing="ij")
    board = np.stack([gx.T.ravel(), gy.T.ravel(), np.zeros(n_pts)], axis=-1) * square
    board -= board.mean(axis=0)
    centres = np.stack([-R[c].T @ t[c] for c in range(C)])
    fwd = np.stack([R[c][2] for c in range(C)])
    pts3d, obs2d, pidx, cidx = [], [], [], []
    views = 0
    tries = 0
    while views < n_views and tries < max_tries_factor * n_views:
        tries += 1
        c0 = int(rng.integers(0, C - 1))
        mid = 0.5 * (centres[c0] + centres[c0 + 1])
        d = fwd[c0] + fwd[c0 + 1]
        d /= np.linalg.norm(d)
        pos = mid + d * rng.uniform(2.0, 6.0) + rng.normal(0, 0.5, 3)
        # board frame: normal roughly facing back at the cameras, random tilt <= 40 deg
        z = -d
        tilt = np.deg2rad(rng.uniform(0, 40.0))
        axis = rng.normal(0, 1, 3)
        axis -= axis.dot(z) * z
        axis /= np.linalg.norm(axis)
        z = np.cos(tilt) * z + np.sin(tilt) * axis
        x = np.cross([0, 0, 1.0], z)
        x /= np.linalg.norm(x)
        roll = rng.uniform(-0.3, 0.3)
        y = np.cross(z, x)
        x, y = np.cos(roll) * x + np.sin(roll) * y, -np.sin(roll) * x + np.cos(roll) * y
        Rb = np.stack([x, y, z], axis=1)
        Xw = board @ Rb.T + pos
        seen = []
        for c in range(C):
            Xc = Xw @ R[c].T + t[c]
            th = np.arctan2(np.hypot(Xc[:, 0], Xc[:, 1]), Xc[:, 2])
            if not np.all((th < np.deg2rad(60.0)) & (Xc[:, 2] > 0)):
                continue
            uv = np.asarray(project(Xw, K[c], D[c], R[c], t[c]), dtype=np.float64)
            if np.all((uv[:, 0] >= 0) & (uv[:, 0] < res[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < res[1])):
                seen.append((c, uv))
        if len(seen) < 2:
            continue
        base = views * n_pts
        pts3d.append(Xw)
        for c, uv in seen:
            obs2d.append(uv + rng.normal(0, noise_px, uv.shape))
            pidx.append(base + np.arange(n_pts))
            cidx.append(np.full(n_pts, c))
        views += 1
    # perturbed extrinsics: rotate by ~rot_sigma about a random axis, shift by trans_sigma
    R0 = np.empty_like(R)
    t0 = np.empty_like(t)
    for c in range(C):
        ax = rng.normal(0, 1, 3)
        ax /= np.linalg.norm(ax)
        a = np.deg2rad(rot_sigma_deg) * rng.normal()
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx
        R0[c] = dR @ R[c]
        t0[c] = t[c] + rng.normal(0, trans_sigma, 3)
    return dict(
        K=K, D=D, R_true=R, t_true=t, R0=R0, t0=t0, res=res,
        points_3d_true=np.concatenate(pts3d), points_2d=np.concatenate(obs2d).astype(np.float32),
        point_3d_indices=np.concatenate(pidx).astype(np.int64),
        camera_indices=np.concatenate(cidx).astype(np.int64), n_views=views, board_shape=board_shape,
    )
