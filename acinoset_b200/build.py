"""The reference's data-driven trajectory builder (/root/reference/src/build.py) by name, on libacino_b200.so.

    load_skeleton(skel_file)                                   build.py:19-26
    build_model(skel_dict, project_dir) -> (model, pose_to_3d) :28-304   (Pyomo ConcreteModel -> SkeletonModel)
    solve_optimisation(model, exe_path, project_dir, poses)    :306-332  (IPOPT subprocess -> GPU Levenberg-Marquardt)
    convert_to_dict / save_data                                :344-378  (same pickle: positions, x, dx, ddx)
    redescending_loss, pt3d_to_2d / _x2d / _y2d, rot_x/y/z     :382-481

What `build_model` keeps from the reference, literally: scene ``data/4_cam_scene_static_sba.json`` and DLC tables
``data/*.h5`` under project_dir (:97,107), h = 1/120, start_frame = 60, N = 100 (:131-133), R = 3 and the 0.4
likelihood threshold (:142,166-170), marker "neck" skipped (:123-124,196-197,281-282), model weight 0.002 for every
parameter (:173-177), initial guess = linear regression of the triangulated "forehead" over the frame number,
evaluated at frames 0..N-1 (:149-157), all angles 0 (:210-216), bounds |x_i| <= pi/2 for 1-based i in [3, 3L) on
frames 1..N-1 (:263-266), objective sum 0.002 slack_model^2 + sum |w slack_meas| (:287-302), and the pairing of FK
output row l with DLC marker ``markers[l]`` BY INDEX (:189-198,219-224,276-285) - for the shipped human skeleton the
two orders differ, which is the reference's behaviour; ``pair_by="name"`` pairs them by name instead.

The Pyomo model object is not a boundary worth keeping: `model` is a plain SkeletonModel that
carries the dense tensors; the solve is a projected Levenberg-Marquardt loop whose every step (evaluation, gradient,
band assembly, band Cholesky, trial point, model reduction, fixed-order sums) is a CUDA kernel (csrc/skel.cu).
There is no CPU path.
"""
import glob
import os
import pickle
from dataclasses import dataclass, field
from time import time

import numpy as np

from . import calib, skeleton, utils
from . import fte as _fte
from .skeleton import load_skeleton  # noqa: F401

from .config import SKELETON  # noqa: E402

MODEL_WEIGHT = SKELETON.model_weight
MEAS_SIGMA_R = SKELETON.meas_sigma_px
LIK_THRESH = SKELETON.lik_thresh


@dataclass
class SkeletonModel:
    """What the reference's ConcreteModel holds, as dense arrays."""
    skel_dict: dict
    flat: dict
    cams: tuple                 # (K, D (C,4), R, t)
    meas: np.ndarray            # (N, C, n_out, 2)
    w: np.ndarray               # (N, C, n_out)   meas_err_weight
    x0: np.ndarray              # (N, P)          initial guess
    h: float = 1.0 / 120.0
    start_frame: int = 60
    loss: str = "abs"
    delta: float = 0.05
    model_weight: float = MODEL_WEIGHT
    last_free: bool = True
    device: int = 0
    x: np.ndarray = None        # solution (after solve_optimisation)
    info: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.meas.shape[0]

    @property
    def P(self):
        return self.x0.shape[1]


def bounds(n_parts):
    """build.py:263-266 (0-based slots 2 .. 3L-2: includes z, excludes the last psi's)."""
    P = 3 + 3 * n_parts
    lo = np.full(P, -np.inf)
    hi = np.full(P, np.inf)
    lo[2:3 * n_parts - 1] = -np.pi / 2
    hi[2:3 * n_parts - 1] = np.pi / 2
    return lo, hi


class SkelSolver:
    """Projected Levenberg-Marquardt on F = sum rho(w r) + sum q (third difference / h^2)^2, one GPU.

    Same algorithm as acinoset_b200.lm.FTESolver (damping lam diag(B), frozen bound-active variables, gain ratio,
    lam /= 3 | *= 2 | *= 4), with the generic kernels: the normal matrix is banded with half bandwidth 3P and is
    factored by one band Cholesky per attempt.  Replicas only across GPUs (SURVEY.md section 8e has no sharding row
    for this variant; its problems are 100 frames)."""

    def __init__(self, handle, flat, meas, w, h, loss="abs", abc=(3.0, 10.0, 20.0), delta=0.05, model_weight=MODEL_WEIGHT,
                 bounds_=None, last_free=True):
        import torch

        self.h, self.torch = handle, torch
        self.dev = torch.device("cuda", handle.device)
        handle.skel_set(flat, loss=loss, abc=abc, delta=delta)
        self.P, self.n_out = handle.skel_shape
        self.N = int(meas.shape[0])
        N, P = self.N, self.P
        f64 = torch.float64
        dev = self.dev
        self.meas = torch.as_tensor(np.ascontiguousarray(meas, dtype=np.float64)).to(dev)
        self.w = torch.as_tensor(np.ascontiguousarray(w, dtype=np.float64)).to(dev)
        lo, hi = bounds((P - 3) // 3) if bounds_ is None else bounds_
        self.lo_np, self.hi_np = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        self.lo = torch.as_tensor(self.lo_np).to(dev)
        self.hi = torch.as_tensor(self.hi_np).to(dev)
        self.sw = torch.full((P,), 2.0 * model_weight / h ** 4, dtype=f64, device=dev)
        self.last_free = 1 if last_free else 0

        def buf(*shape, dtype=f64):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        NU = P * (P + 1) // 2
        self.st = [dict(x=buf(N, P), cost=buf(N), g=buf(N, P), H=buf(N, NU), gtot=buf(N, P),
                        fixed=buf(N, P, dtype=torch.uint8), cost_s=buf(N)) for _ in range(2)]
        self.AB = buf(N * P, 3 * P + 1)
        self.d = buf(N * P)
        self.pred, self.step = buf(N), buf(N)
        self.info = buf(1, dtype=torch.int32)
        self.out5 = buf(5)
        self.n_launch0 = handle.launch_count

    def _eval(self, s):
        self.h.call_dev("acino_skel_eval_dev", self.N, s["x"], self.meas, self.w, s["cost"], s["g"], s["H"])
        self.h.call_dev("acino_skel_prepare_dev", self.N, self.last_free, s["x"], s["g"], self.sw, self.lo, self.hi,
                        s["gtot"], s["fixed"], s["cost_s"])

    def _sums(self, s, with_step):
        self.h.call_dev("acino_lm_reduce_dev", self.N, None, s["cost"], s["cost_s"], self.pred if with_step else None,
                        self.step if with_step else None, self.out5)
        o = self.out5.cpu().numpy()
        return float(o[1] + o[2]), float(o[3]), float(o[4])

    def solve(self, x0, max_iter=100, lam0=1e-3, tol_step=1e-6, tol_rel=1e-8, max_attempts=12, verbose=False):
        torch = self.torch
        N, P = self.N, self.P
        s, t = self.st
        x0 = np.array(x0, dtype=np.float64)
        nb = N - 1 if self.last_free else N
        x0[:nb] = np.clip(x0[:nb], self.lo_np, self.hi_np)
        s["x"].copy_(torch.as_tensor(x0).to(self.dev))
        self._eval(s)
        F, _, _ = self._sums(s, with_step=False)
        lam, hist = lam0, [F]
        n_solve, it, converged, bad_pivot = 0, 0, False, 0
        for it in range(max_iter):
            accepted = False
            for _ in range(max_attempts):
                self.h.call_dev("acino_skel_assemble_dev", N, s["H"], s["gtot"], s["fixed"], self.sw, float(lam), self.AB, self.d)
                self.info.zero_()
                self.h.call_dev("acino_band_solve_dev", N * P, 3 * P, self.AB, self.d, self.info)
                n_solve += 1
                self.h.call_dev("acino_skel_trial_dev", N, self.last_free, s["x"], self.d, self.lo, self.hi, t["x"])
                self.h.call_dev("acino_skel_pred_dev", N, s["x"], t["x"], s["gtot"], s["H"], self.sw, self.pred, self.step)
                self._eval(t)
                Ft, pred, step = self._sums(t, with_step=True)
                if int(self.info.item()) != 0:        # not positive definite at this damping: raise it
                    bad_pivot += 1
                    lam *= 4
                    continue
                rho = (F - Ft) / pred if pred > 0 else -1.0
                if verbose:
                    print(f"it {it:3d} lam {lam:9.3e} F {F:16.6f} Ft {Ft:16.6f} pred {pred:10.3e} rho {rho:7.3f} |dx|inf {step:.2e}")
                if Ft < F and rho > 1e-4:
                    accepted = True
                    rel = (F - Ft) / max(abs(F), 1e-30)
                    F = Ft
                    s, t = t, s
                    lam = max(lam / 3, 1e-12) if rho > 0.75 else (lam * 2 if rho < 0.25 else lam)
                    break
                lam *= 4
            hist.append(F)
            if not accepted:
                break
            if step < tol_step or rel < tol_rel:
                converged = True
                break
        self.st = [s, t]
        torch.cuda.synchronize(self.dev)
        info = dict(F=F, F0=hist[0], iters=it + 1, n_solve=n_solve, history=hist, lam=lam, converged=converged,
                    bad_pivots=bad_pivot, launches=self.h.launch_count - self.n_launch0)
        return s["x"].cpu().numpy(), info


def model_from_arrays(skel_dict, cams, meas, lik, marker_names=None, x0=None, h=1.0 / 120.0, start_frame=0, loss="abs",
                      delta=0.05, pair_by="index", device=0):
    """The array-level half of build_model: cams = (K, D, R, t); meas (N, C, n_markers, 2) / lik (N, C, n_markers) in
    the order of ``marker_names`` (default: skel_dict["markers"], or the FK row order if that list is empty)."""
    flat = skeleton.flatten_skeleton(skel_dict)
    out_names = flat["out_names"]
    n_out = len(out_names)
    markers = list(marker_names) if marker_names is not None else (list(skel_dict["markers"]) or list(out_names))
    meas = np.asarray(meas, dtype=np.float64)
    lik = np.asarray(lik, dtype=np.float64)
    N, C = meas.shape[0], meas.shape[1]
    m_out = np.zeros((N, C, n_out, 2))
    w_out = np.zeros((N, C, n_out))
    for r in range(n_out):
        if pair_by == "index":          # build.py: pos_funcs[l-1] <-> markers[l-1]
            j = r if r < len(markers) else None
        else:
            j = markers.index(out_names[r]) if out_names[r] in markers else None
        if j is None or markers[j] == "neck":      # build.py:123-124,196-197,281-282
            continue
        m_out[:, :, r] = meas[:, :, j]
        w_out[:, :, r] = np.where(lik[:, :, j] > LIK_THRESH, 1.0 / MEAS_SIGMA_R, 0.0)
    K, D, R, t = cams
    D = np.asarray(D, dtype=np.float64).reshape(-1, 4)
    P = 3 + 3 * len(flat["parts"])
    x0 = np.zeros((N, P)) if x0 is None else np.asarray(x0, dtype=np.float64)
    return SkeletonModel(skel_dict=skel_dict, flat=flat, cams=(K, D, R, t), meas=m_out, w=w_out, x0=x0, h=h,
                         start_frame=start_frame, loss=loss, delta=delta, device=device)


def build_model(skel_dict, project_dir, N=100, start_frame=60, h=1.0 / 120.0, pair_by="index", device=0):
    """build.py:28-304 -> (model, pose_to_3d)."""
    from scipy.stats import linregress

    scene_path = os.path.join(project_dir, "data", "4_cam_scene_static_sba.json")
    K_arr, D_arr, R_arr, t_arr, _ = utils.load_scene(scene_path)
    D_arr = D_arr.reshape((-1, 4))
    print("\n\n\nLoading data")
    df_paths = sorted(glob.glob(os.path.join(project_dir, "data", "*.h5"))) or \
        sorted(glob.glob(os.path.join(project_dir, "data", "*.csv")))
    points_2d_df = utils.create_dlc_points_2d_file(df_paths, verbose=False) if df_paths else None
    if points_2d_df is None or not len(points_2d_df):
        raise FileNotFoundError(f"no DLC tables under {os.path.join(project_dir, 'data')}")
    markers = list(skel_dict["markers"])
    C = len(K_arr)
    # frames n = 1..N read DLC frame n + start_frame - 1 (get_meas_from_df, build.py:111-128)
    meas, lik = utils.dlc_df_to_dense(points_2d_df, C, markers, start_frame, N)
    # initial guess (build.py:144-157,210-216)
    pts3d = calib.get_pairwise_3d_points_from_df(points_2d_df[points_2d_df["likelihood"] > LIK_THRESH], K_arr, D_arr, R_arr,
                                                 t_arr, calib.triangulate_points_fisheye, device=device)
    nose = pts3d[pts3d["marker"] == "forehead"][["x", "y", "z", "frame"]].to_numpy(dtype=np.float64)
    P = 3 + 3 * len(skel_dict["positions"])
    x0 = np.zeros((N, P))
    if len(nose) >= 2:
        fr = np.arange(N)
        for k in range(3):
            lr = linregress(nose[:, 3], nose[:, k])
            x0[:, k] = fr * lr.slope + lr.intercept
    model = model_from_arrays(skel_dict, (K_arr, D_arr, R_arr, t_arr), meas, lik, markers, x0, h, start_frame,
                              pair_by=pair_by, device=device)
    return model, skeleton.build_pose_function(skel_dict, device=device)


def solve_optimisation(model, exe_path=None, project_dir=None, poses=None, verbose=False, **lm_kwargs):
    """build.py:306-332.  ``exe_path`` (the IPOPT binary) is accepted and ignored.  Writes
    ``project_dir/data/results/traj_results.pickle`` when project_dir is given; returns the result dict."""
    t0 = time()
    K, D, R, t = model.cams
    handle = _fte.set_scene(K, D, R, t, model.device)
    solver = SkelSolver(handle, model.flat, model.meas, model.w, model.h, loss=model.loss, delta=model.delta,
                        model_weight=model.model_weight, last_free=model.last_free)
    model.x, model.info = solver.solve(model.x0, verbose=verbose, **lm_kwargs)
    print("Optimization took {0:.2f} seconds".format(time() - t0))
    poses = skeleton.build_pose_function(model.skel_dict, device=model.device) if poses is None else poses
    if project_dir is not None:
        save_data(model, os.path.join(project_dir, "data", "results", "traj_results.pickle"), poses)
    return convert_to_dict(model, poses)


def convert_to_dict(m, poses):
    """build.py:344-366: {positions, x, dx, ddx}; dx / ddx follow the collocation constraints (:231-261)."""
    x = np.asarray(m.x, dtype=np.float64)
    dx, ddx = _fte.derived_velocities(x, m.h)
    positions = np.asarray(poses(x))
    return dict(positions=positions, x=x, dx=dx, ddx=ddx)


def save_data(file_data, file_path, poses, dict=True):
    """build.py:368-378"""
    if dict:
        file_data = convert_to_dict(file_data, poses)
    os.makedirs(os.path.dirname(file_path), exist_ok=True)
    with open(file_path, "wb") as f:
        pickle.dump(file_data, f)
    print(f"save {file_path}")


# ---- helpers the reference module also exports -------------------------------------------------------------
def func_step(start, x):
    """build.py:382-383"""
    return 1 / (1 + np.e ** (-1 * (x - start)))


def func_piece(start, end, x):
    """build.py:385-386"""
    return func_step(start, x) - func_step(end, x)


def redescending_loss(err, a, b, c):
    """build.py:388-395 (host scalar / array form of the loss the kernels evaluate)."""
    e = abs(err)
    cost = (1 - func_step(a, e)) / 2 * e ** 2
    cost += func_piece(a, b, e) * (a * e - (a ** 2) / 2)
    cost += func_piece(b, c, e) * (a * b - (a ** 2) / 2 + (a * (c - b) / 2) * (1 - ((c - e) / (c - b)) ** 2))
    cost += func_step(c, e) * (a * b - (a ** 2) / 2 + (a * (c - b) / 2))
    return cost


def np_rot_x(x):
    """build.py:428-434"""
    c, s = np.cos(x), np.sin(x)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])


def np_rot_y(y):
    """build.py:436-443"""
    c, s = np.cos(y), np.sin(y)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])


def np_rot_z(z):
    """build.py:445-452"""
    c, s = np.cos(z), np.sin(z)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])


def pt3d_to_2d(x, y, z, K, D, R, t, device=0):
    """build.py:457-473"""
    from . import all_optimizations

    return all_optimizations.pt3d_to_2d(x, y, z, K, D, R, t, device)


def pt3d_to_x2d(x, y, z, K, D, R, t, device=0):
    """build.py:475-477"""
    return pt3d_to_2d(x, y, z, K, D, R, t, device)[0]


def pt3d_to_y2d(x, y, z, K, D, R, t, device=0):
    """build.py:479-481"""
    return pt3d_to_2d(x, y, z, K, D, R, t, device)[1]
