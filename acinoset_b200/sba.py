"""Sparse bundle adjustment on the GPU behind the reference's entry points
(/root/reference/src/calib/calib.py:196-390 and app.py:201-223).

Same names, argument order, shapes and return values as the reference:
    create_bundle_adjustment_jacobian_sparsity_matrix   calib.py:196-207
    prepare_calib_board_data_for_bundle_adjustment      calib.py:210-263
    prepare_manual_points_for_bundle_adjustment         calib.py:266-304
    params_to_points_only / cost_func_points_only       calib.py:307-316
    params_to_points_extrinsics / cost_func_points_extrinsics   calib.py:345-359
    bundle_adjust_points_only / bundle_adjust_board_points_only calib.py:319-341
    bundle_adjust_points_and_extrinsics / bundle_adjust_board_points_and_extrinsics  calib.py:362-390
    sba_board_points_fisheye                             app.py:201-223
The reference minimises  0.5 sum C^2 ln(1 + (f/C)^2)  (SciPy least_squares, loss='cauchy',
f_scale=C) with a finite-difference Jacobian over one OpenCV call per observation; here the
residuals, analytic Jacobian blocks, the per-point Schur complement and the 6C x 6C solve run in
libacino_b200.so (csrc/sba.cu) inside a Levenberg-Marquardt loop.  ``project_func`` /
``triangulate_func`` select the camera model like in the reference: the fisheye pair
(``project_points_fisheye`` / ``triangulate_points_fisheye``, also the default) or the standard-model pair
(``project_points`` / ``triangulate_points``: what ``app.sba_board_points`` passes, app.py:215-218); any other callable
raises - there is no per-observation host callback path.
"""
import time

import numpy as np

from . import calib as _calib
from . import fte as _fte
from . import _lib


from .rotations import rodrigues_to_mat, rodrigues_to_vec  # noqa: E402,F401


# ---- camera model dispatch -------------------------------------------------------------------------------
FISHEYE, PINHOLE = 0, 1


def camera_model(project_func=None, triangulate_func=None):
    """Camera model the reference's function arguments stand for: FISHEYE (k1..k4, cv2.fisheye) for
    ``*_fisheye`` or None, PINHOLE (OpenCV's standard model, up to 12 coefficients) for ``project_points`` /
    ``triangulate_points``.  Matching is by function name, so the reference's own functions select the same model."""
    names = {getattr(f, "__name__", repr(f)) for f in (project_func, triangulate_func) if f is not None}
    fish = {"project_points_fisheye", "triangulate_points_fisheye"}
    pin = {"project_points", "triangulate_points"}
    if names <= fish:
        return FISHEYE
    if names <= pin:
        return PINHOLE
    raise NotImplementedError(
        f"camera functions {sorted(names)}: the GPU bundle adjustment carries the fisheye and the standard OpenCV model "
        "(calib.project_points[_fisheye] / triangulate_points[_fisheye]); pass one of those pairs")


def _triangulate(model):
    return _calib.triangulate_points if model == PINHOLE else _calib.triangulate_points_fisheye


def _dist_table(d_arr, n_cams, model):
    """(C, n_dist) fp64 coefficient table: 4 fisheye coefficients, or the standard model's up to 12
    [k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4] zero-padded to the longest camera (tilt terms must be zero)."""
    if model == FISHEYE:
        return np.array([np.asarray(d, dtype=np.float64).reshape(-1)[:4] for d in d_arr]).reshape(n_cams, 4)
    ds = [np.zeros(0) if d is None else np.asarray(d, dtype=np.float64).reshape(-1) for d in d_arr]
    for d in ds:
        if d.size > 12 and np.any(d[12:] != 0):
            raise ValueError("tilted-sensor distortion terms (tauX, tauY) are not supported")
    nd = max(4, min(12, max(d.size for d in ds)))
    out = np.zeros((n_cams, nd))
    for i, d in enumerate(ds):
        out[i, :min(d.size, nd)] = d[:nd]
    return out


# ---- problem assembly (host) -----------------------------------------------------------------------
def create_bundle_adjustment_jacobian_sparsity_matrix(n_cameras, n_params_per_camera, camera_indices, n_points,
                                                      point_indices):
    """Sparsity pattern with the reference's shape and column layout (calib.py:196-207): rows 2i, 2i+1 of observation i
    touch the ``n_params_per_camera`` columns of a CAMERA-MAJOR block ``camera * n_params_per_camera + s`` and the three
    columns of its point.  (The reference's parameter vector is [all rvecs | all tvecs | points], calib.py:346-351, so
    for the extrinsics problem this pattern does not line up with it; kept for callers that pass it on - the solves
    here use the analytic Jacobian, ``jac_points_extrinsics``.)  Built in one shot as a COO matrix."""
    from scipy.sparse import coo_matrix

    cam = np.asarray(camera_indices, dtype=np.int64).ravel()
    pt = np.asarray(point_indices, dtype=np.int64).ravel()
    n_obs, npc = cam.size, int(n_params_per_camera)
    cols = np.concatenate([cam[:, None] * npc + np.arange(npc)[None, :],
                           n_cameras * npc + pt[:, None] * 3 + np.arange(3)[None, :]], axis=1)        # (n_obs, npc + 3)
    rows = 2 * np.arange(n_obs)[:, None, None] + np.arange(2)[None, :, None]                               # (n_obs, 2, 1)
    rows = np.broadcast_to(rows, (n_obs, 2, npc + 3))
    cols = np.broadcast_to(cols[:, None, :], (n_obs, 2, npc + 3))
    A = coo_matrix((np.ones(rows.size, dtype=int), (rows.ravel(), cols.ravel())),
                   shape=(2 * n_obs, n_cameras * npc + n_points * 3))
    return A.tolil()


def prepare_calib_board_data_for_bundle_adjustment(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr, r_arr, t_arr,
                                                   triangulate_func=None):
    """calib.py:210-263.  Views (file names) seen by >= 2 cameras; 54 new points per view, initial
    estimate triangulated from the FIRST TWO cameras that see the view.  All views of one camera pair
    are triangulated by a single kernel launch.  View order = sorted names (the reference iterates a
    Python set; the cost is order-invariant)."""
    tri = _triangulate(camera_model(None, triangulate_func))
    n_cam = len(img_pts_arr)
    fnames_arr = [list(f) for f in fnames_arr]
    counts = {}
    for fnames in fnames_arr:
        for f in fnames:
            counts[f] = counts.get(f, 0) + 1
    views = sorted(f for f, v in counts.items() if v >= 2)
    ppi = board_shape[0] * board_shape[1]
    lookup = [{f: i for i, f in enumerate(fn)} for fn in fnames_arr]
    points_2d, point_3d_indices, camera_indices = [], [], []
    pair_views = {}
    for v, fname in enumerate(views):
        seen = [c for c in range(n_cam) if fname in lookup[c]]
        for c in seen:
            points_2d.append(np.asarray(img_pts_arr[c][lookup[c][fname]], dtype=np.float32).reshape(ppi, 2))
            point_3d_indices.append(np.arange(v * ppi, (v + 1) * ppi))
            camera_indices.append(np.full(ppi, c))
        pair_views.setdefault((seen[0], seen[1]), []).append(v)
    points_3d = np.zeros((len(views) * ppi, 3), dtype=np.float64)
    for (a, b), vs in pair_views.items():
        pa = np.concatenate([np.asarray(img_pts_arr[a][lookup[a][views[v]]], dtype=np.float64).reshape(ppi, 2) for v in vs])
        pb = np.concatenate([np.asarray(img_pts_arr[b][lookup[b][views[v]]], dtype=np.float64).reshape(ppi, 2) for v in vs])
        X = tri(pa, pb, k_arr[a], d_arr[a], r_arr[a], t_arr[a], k_arr[b], d_arr[b], r_arr[b], t_arr[b])
        for j, v in enumerate(vs):
            points_3d[v * ppi:(v + 1) * ppi] = X[j * ppi:(j + 1) * ppi]
    if not views:
        return (np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64))
    return (np.concatenate(points_2d).astype(np.float32), points_3d.astype(np.float32),
            np.concatenate(point_3d_indices).astype(np.int64), np.concatenate(camera_indices).astype(np.int64))


def prepare_manual_points_for_bundle_adjustment(img_pts_arr, k_arr, d_arr, r_arr, t_arr, triangulate_func=None):
    """calib.py:266-304: img_pts_arr (n_points, n_cameras, 2) with NaN where unseen."""
    tri = _triangulate(camera_model(None, triangulate_func))
    pts = np.asarray(img_pts_arr, dtype=np.float64).swapaxes(0, 1)
    n_cam, n_pts = pts.shape[0], pts.shape[1]
    seen = ~np.isnan(pts).any(axis=2)                       # (n_cam, n_pts)
    keep = np.nonzero(seen.sum(axis=0) > 1)[0]
    points_2d, point_3d_indices, camera_indices = [], [], []
    first_two = {}
    for new_idx, i in enumerate(keep):
        cams = np.nonzero(seen[:, i])[0]
        for c in cams:
            points_2d.append(pts[c, i])
            camera_indices.append(c)
            point_3d_indices.append(new_idx)
        first_two.setdefault((cams[0], cams[1]), []).append((new_idx, i))
    points_3d = np.zeros((len(keep), 3))
    for (a, b), lst in first_two.items():
        idx_new = [x[0] for x in lst]
        idx_old = [x[1] for x in lst]
        X = tri(pts[a, idx_old], pts[b, idx_old], k_arr[a], d_arr[a], r_arr[a], t_arr[a], k_arr[b], d_arr[b], r_arr[b],
                t_arr[b])
        points_3d[idx_new] = X
    return (np.array(points_2d, dtype=np.float32).reshape(-1, 2), points_3d.astype(np.float32),
            np.array(point_3d_indices, dtype=np.int64), np.array(camera_indices, dtype=np.int64))


def params_to_points_only(params, n_points):
    return np.asarray(params).reshape((n_points, 3))


def params_to_points_extrinsics(params, n_cameras, n_points):
    params = np.asarray(params, dtype=np.float64)
    r_end = n_cameras * 3
    t_end = r_end + n_cameras * 3
    r_arr = np.array([rodrigues_to_mat(r) for r in params[:r_end].reshape((n_cameras, 3))], dtype=np.float64)
    t_arr = params[r_end:t_end].reshape((n_cameras, 3, 1))
    obj_pts = params[t_end:].reshape((n_points, 3))
    return obj_pts, r_arr, t_arr



# ---- multi-GPU plan (pure NumPy / torch: runs on CPU tensors with gloo in the tests) ----------------
def shard_points(point_3d_indices, n_points, world):
    """Shard the 3-D points (board views) over ranks in contiguous index ranges; every point's
    observations follow it, so the point blocks stay private to a rank and only the 6C camera
    parameters are shared.  Returns a list of
    (point0, n_points_local, obs_ids) with obs_ids the global observation ids of the shard."""
    pidx = np.asarray(point_3d_indices, dtype=np.int64)
    out = []
    for r in range(world):
        p0 = (r * n_points) // world
        p1 = ((r + 1) * n_points) // world
        ids = np.nonzero((pidx >= p0) & (pidx < p1))[0]
        out.append((p0, p1 - p0, ids))
    return out


def allreduce_camera_system(Sr, world, group=None):
    """The ONE data-path collective of a sharded SBA iteration: sum of the per-rank reduced camera
    systems, packed [S (6C x 6C) | rhs (6C)] in one buffer (5.4 KB at C = 6).  Every rank receives
    the same bits, so the redundant dense solves that follow agree exactly."""
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(Sr, op=dist.ReduceOp.SUM, group=group)
    return Sr


def allreduce_scalars(v, world, group=None):
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v


def scatter_sum(local, ids, n_total, world, group=None):
    """Place a rank's rows at their global ids and sum over ranks (the final gather of the sharded
    points / residuals; done once per solve, outside the iteration)."""
    import torch

    full = torch.zeros((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    full[ids] = local
    return allreduce_scalars(full, world, group)

# ---- device problem ----------------------------------------------------------------------------------
class SBAProblem:
    """Observations + cameras resident on one GPU; evaluates residuals / Jacobian blocks and runs LM."""

    def __init__(self, points_2d, point_3d_indices, camera_indices, k_arr, d_arr, n_points, device=0,
                 with_extrinsics=True, f_scale=1.0, rank=0, world=1, group=None, model=FISHEYE):
        import torch

        self.torch = torch
        self.rank, self.world, self.group = int(rank), int(world), group
        self.h = _fte.get_handle(device)
        self.dev = torch.device("cuda", device)
        self.C = len(k_arr)
        if self.C > 10:
            raise ValueError("SBA supports up to 10 cameras")
        self.n_obs = int(len(point_3d_indices))
        self.n_pts = int(n_points)
        self.with_ext = with_extrinsics
        self.f_scale = float(f_scale)
        dev, f64 = self.dev, torch.float64
        pidx = np.asarray(point_3d_indices, dtype=np.int64)
        cidx = np.asarray(camera_indices, dtype=np.int64)
        self.uv = torch.as_tensor(np.ascontiguousarray(points_2d, dtype=np.float32).reshape(-1, 2)).to(dev)
        self.cam_idx = torch.as_tensor(cidx.astype(np.int32)).to(dev)
        self.pt_idx = torch.as_tensor(pidx.astype(np.int32)).to(dev)
        order = np.argsort(pidx, kind="stable").astype(np.int32)        # CSR by point, cameras ascending
        ptr = np.zeros(self.n_pts + 1, dtype=np.int32)
        np.cumsum(np.bincount(pidx, minlength=self.n_pts), out=ptr[1:])
        self.obs = torch.as_tensor(order).to(dev)
        self.pt_ptr = torch.as_tensor(ptr).to(dev)
        self.K = torch.as_tensor(np.asarray(k_arr, dtype=np.float64).reshape(self.C, 9)).to(dev)
        self.model = int(model)
        dtab = _dist_table(d_arr, self.C, self.model)
        self.n_dist = dtab.shape[1]
        self.D = torch.as_tensor(dtab).to(dev)
        cam_bytes = int(_lib.lib.acino_sba_cam_bytes())
        self.cams = [torch.zeros(self.C * cam_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        n = self.n_obs

        def buf(*shape):
            return torch.zeros(*shape, dtype=f64, device=dev)

        self.st = [dict(params=buf(6 * self.C), pts=buf(self.n_pts, 3), res=buf(n, 2), Jc=buf(n, 2, 6), Jp=buf(n, 2, 3),
                        wgt=buf(n, 2), cost=buf(n), cams=self.cams[i]) for i in range(2)]
        self.pred = buf(n)
        self.dp = buf(self.n_pts, 3)
        n6 = 6 * self.C
        self.Sr = buf(n6 * n6 + n6)                 # [S | rhs]: one buffer = one all_reduce when sharded
        self.S = self.Sr[:n6 * n6].view(n6, n6)
        self.dc = self.Sr[n6 * n6:]
        self.partial = buf(int(_lib.lib.acino_sba_schur_partial_size(self.n_pts, self.C)))
        self.scal = buf(9)                          # [lm_reduce out5 | step / state norms]: ONE device->host copy per attempt
        self.out5 = self.scal[:5]
        self.sc4 = self.scal[5:]
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self.Rfix = self.tfix = None
        # pinned host mirrors of the solve's results (residuals before / after: 16 bytes per observation - pageable
        # copies of these were two thirds of the wall time of a configs[3] solve)
        self._host = None                           # allocated by the first solve (residuals() / jacobian_blocks() need none)

    def set_fixed_cameras(self, r_arr, t_arr):
        t_ = self.torch
        # the reference's project_func converts every r to a Rodrigues vector and back (calib.py:134; cv2.projectPoints
        # does the same with a matrix argument): a slightly non-orthonormal scene matrix is projected onto SO(3) exactly
        # like cv2.Rodrigues does
        r_arr = np.array([rodrigues_to_mat(rodrigues_to_vec(r)) for r in np.asarray(r_arr, dtype=np.float64).reshape(-1, 3, 3)])
        self.Rfix = t_.as_tensor(np.asarray(r_arr, dtype=np.float64).reshape(self.C, 9)).to(self.dev)
        self.tfix = t_.as_tensor(np.asarray(t_arr, dtype=np.float64).reshape(self.C, 3)).to(self.dev)

    # -- evaluation
    def _eval(self, s, want_j=True):
        if self.with_ext:
            self.h.call_dev("acino_sba_cams_model_dev", self.C, self.model, self.n_dist, s["params"], None, None, self.K,
                            self.D, s["cams"])
        else:
            self.h.call_dev("acino_sba_cams_model_dev", self.C, self.model, self.n_dist, None, self.Rfix, self.tfix, self.K,
                            self.D, s["cams"])
        self.h.call_dev("acino_sba_eval_dev", self.n_obs, s["cams"], s["pts"], self.uv, self.cam_idx, self.pt_idx,
                        self.f_scale, s["res"], (s["Jc"] if self.with_ext else None) if want_j else None,
                        s["Jp"] if want_j else None, s["wgt"] if want_j else None, s["cost"])

    def _sum(self, a1, a2=None, norms_of=None):
        """Fixed-order local sums (+ one small all_reduce when the points are sharded) -> (sum a1, sum a2[, |dp|^2,
        |pts|^2, |dc|^2, |params|^2 of the trial state ``norms_of``]) with a single device->host copy."""
        self.h.call_dev("acino_lm_reduce_dev", self.n_obs, None, a1, a2, None, None, self.out5)
        if norms_of is not None:
            t_ = self.torch
            t_.sum(self.dp * self.dp, dim=None, out=self.sc4[0])
            t_.sum(norms_of["pts"] * norms_of["pts"], dim=None, out=self.sc4[1])
            if self.with_ext:
                t_.sum(self.dc * self.dc, dim=None, out=self.sc4[2])
                t_.sum(norms_of["params"] * norms_of["params"], dim=None, out=self.sc4[3])
            if self.world > 1:
                self._allreduce_scal()
        else:
            allreduce_scalars(self.out5[1:3], self.world, self.group)
        o = self.scal.cpu().numpy()
        return (float(o[1]), float(o[2])) if norms_of is None else (float(o[1]), float(o[2]), o[5:9].copy())

    def _allreduce_scal(self):
        """cost / pred / point-step / point-state sums are sharded over the ranks; the camera norms are replicated."""
        v = self.scal[[1, 2, 5, 6]]
        allreduce_scalars(v, self.world, self.group)
        self.scal[[1, 2, 5, 6]] = v

    def residuals(self, params_c, pts):
        """f (2 n_obs,) at the given parameters (no Jacobian)."""
        t_ = self.torch
        s = self.st[1]
        if self.with_ext:
            s["params"].copy_(t_.as_tensor(np.asarray(params_c, dtype=np.float64)).to(self.dev))
        s["pts"].copy_(t_.as_tensor(np.asarray(pts, dtype=np.float64).reshape(-1, 3)).to(self.dev))
        self._eval(s, want_j=False)
        return s["res"].cpu().numpy().ravel()

    def jacobian_blocks(self, params_c, pts):
        """(res (n,2), Jc (n,2,6) or None, Jp (n,2,3), wgt (n,2)) at the given parameters."""
        t_ = self.torch
        s = self.st[1]
        if self.with_ext:
            s["params"].copy_(t_.as_tensor(np.asarray(params_c, dtype=np.float64)).to(self.dev))
        s["pts"].copy_(t_.as_tensor(np.asarray(pts, dtype=np.float64).reshape(-1, 3)).to(self.dev))
        self._eval(s, want_j=True)
        return (s["res"].cpu().numpy(), s["Jc"].cpu().numpy() if self.with_ext else None, s["Jp"].cpu().numpy(),
                s["wgt"].cpu().numpy())

    # -- LM
    def solve(self, params_c0, pts0, max_nfev=1000, ftol=1e-10, xtol=1e-8, lam0=1e-3, verbose=0):
        t_ = self.torch
        s, t = self.st
        if self._host is None:
            f64 = t_.float64
            self._host = dict(f0=t_.empty((self.n_obs, 2), dtype=f64).pin_memory(),
                              fun=t_.empty((self.n_obs, 2), dtype=f64).pin_memory(),
                              pts=t_.empty((self.n_pts, 3), dtype=f64).pin_memory(),
                              params=t_.empty(6 * self.C, dtype=f64).pin_memory())
        if self.with_ext:
            s["params"].copy_(t_.as_tensor(np.asarray(params_c0, dtype=np.float64)).to(self.dev))
        s["pts"].copy_(t_.as_tensor(np.asarray(pts0, dtype=np.float64).reshape(-1, 3)).to(self.dev))
        self._eval(s)
        self._host["f0"].copy_(s["res"], non_blocking=True)       # lands before the first host sync below
        F, _ = self._sum(s["cost"])
        F0 = F
        lam = lam0
        nfev, it = 1, 0
        status = "max_nfev"
        n6 = 6 * self.C
        while nfev < max_nfev:
            it += 1
            accepted = False
            for _ in range(15):
                if self.with_ext:
                    self.h.call_dev("acino_sba_schur_dev", self.n_pts, self.C, self.pt_ptr, self.obs, self.cam_idx, s["res"],
                                    s["Jc"], s["Jp"], s["wgt"], float(lam), self.partial, self.S, self.dc)
                    allreduce_camera_system(self.Sr, self.world, self.group)
                    self.h.call_dev("acino_sba_dense_solve_dev", n6, self.S, self.dc, self.info)
                self.h.call_dev("acino_sba_backsub_dev", self.n_pts, self.C, self.pt_ptr, self.obs, self.cam_idx, s["res"],
                                s["Jc"] if self.with_ext else None, s["Jp"], s["wgt"], float(lam),
                                self.dc if self.with_ext else None, s["pts"], t["pts"], self.dp)
                self.h.call_dev("acino_sba_pred_dev", self.n_obs, self.cam_idx, self.pt_idx, s["res"],
                                s["Jc"] if self.with_ext else None, s["Jp"], s["wgt"], self.dc if self.with_ext else None,
                                self.dp, self.pred)
                if self.with_ext:
                    # dc is camera-major [rvec_a | t_a]; the parameter vector is [all rvecs | all t]
                    t_.add(s["params"], self.dc.view(self.C, 2, 3).transpose(0, 1).reshape(-1), out=t["params"])
                self._eval(t)
                nfev += 1
                Ft, pred, nrm = self._sum(t["cost"], self.pred, norms_of=t)
                rho = (F - Ft) / pred if pred > 0 else -1.0
                if verbose >= 2:
                    print(f"   it {it:4d} nfev {nfev:4d} lam {lam:9.2e} cost {F:.6e} -> {Ft:.6e} pred {pred:9.2e} rho {rho:6.3f}")
                if Ft < F and rho > 1e-4:
                    accepted = True
                    dF = F - Ft
                    xs = float(np.sqrt(nrm[0] + (nrm[2] if self.with_ext else 0.0)))
                    xn = float(np.sqrt(nrm[1] + (nrm[3] if self.with_ext else 0.0)))
                    F = Ft
                    s, t = t, s
                    lam = max(lam / 3, 1e-15) if rho > 0.75 else (lam * 2 if rho < 0.25 else lam)
                    break
                lam *= 4
                if nfev >= max_nfev:
                    break
            if not accepted:
                status = "no_progress" if nfev < max_nfev else "max_nfev"
                break
            if dF < ftol * F:
                status = "ftol"
                break
            if xs < xtol * (xtol + xn):
                status = "xtol"
                break
        self.st = [s, t]
        hb = self._host
        hb["params"].copy_(s["params"], non_blocking=True)
        hb["pts"].copy_(s["pts"], non_blocking=True)
        hb["fun"].copy_(s["res"], non_blocking=True)
        info = int(self.info.item())                                # (synchronises the stream: the copies above are done)
        t_.cuda.synchronize(self.dev)
        # the arrays handed out ARE the pinned mirrors (no second host copy of 16 bytes per observation): they stay valid
        # until the next solve() of this object overwrites them - the reference-named entry points build one object per call
        return dict(params=hb["params"].numpy().copy(), pts=hb["pts"].numpy(), fun=hb["fun"].numpy().ravel(),
                    f0=hb["f0"].numpy().ravel(), cost0=F0, cost=F, nfev=nfev, iters=it, status=status, info=info)


# ---- reference-named entry points ------------------------------------------------------------------
def _dist_ctx():
    """(rank, world, device): a torchrun launch with an initialised process group shards the points
    over the ranks; a plain call is the single-GPU solve."""
    import os

    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", 0))
    except ImportError:
        pass
    return 0, 1, 0


def _solve_sharded(points_2d, points_3d, point_3d_indices, camera_indices, k_arr, d_arr, x_cam0, fixed_rt, f_scale,
                   max_nfev, ftol, verbose, model=FISHEYE, device=None):
    """Common driver: shard the points over the ranks of the process group (if any), solve, and
    return the full-size result on every rank.  ``device``: CUDA device index; None = LOCAL_RANK under a process
    group, else 0."""
    import torch

    rank, world, dev_ctx = _dist_ctx()
    device = dev_ctx if device is None else int(device)
    n_points = len(points_3d)
    pidx = np.asarray(point_3d_indices, dtype=np.int64)
    cidx = np.asarray(camera_indices, dtype=np.int64)
    p2 = np.asarray(points_2d, dtype=np.float32).reshape(-1, 2)
    p0, npl, ids = shard_points(pidx, n_points, world)[rank]
    prob = SBAProblem(p2[ids], pidx[ids] - p0, cidx[ids], k_arr, d_arr, npl, device=device,
                      with_extrinsics=fixed_rt is None, f_scale=f_scale, rank=rank, world=world, model=model)
    if fixed_rt is not None:
        prob.set_fixed_cameras(*fixed_rt)
    out = prob.solve(x_cam0, np.asarray(points_3d, dtype=np.float64)[p0:p0 + npl], max_nfev=max_nfev, ftol=ftol,
                     verbose=verbose if rank == 0 else 0)
    if world > 1:
        dev = prob.dev
        ids_t = torch.as_tensor(ids).to(dev)
        pts = scatter_sum(torch.as_tensor(out["pts"]).to(dev), torch.arange(p0, p0 + npl, device=dev), n_points, world)
        fun = scatter_sum(torch.as_tensor(out["fun"].reshape(-1, 2)).to(dev), ids_t, len(pidx), world)
        f0 = scatter_sum(torch.as_tensor(out["f0"].reshape(-1, 2)).to(dev), ids_t, len(pidx), world)
        out = dict(out, pts=pts.cpu().numpy(), fun=fun.cpu().numpy().ravel(), f0=f0.cpu().numpy().ravel())
    out["world"] = world
    return out


def cost_func_points_only(params, n_points, point_3d_indices, camera_indices, k_arr, d_arr, r_arr, t_arr, points_2d,
                          project_func=None, device=0):
    prob = SBAProblem(points_2d, point_3d_indices, camera_indices, k_arr, d_arr, n_points, with_extrinsics=False,
                      model=camera_model(project_func), device=device)
    prob.set_fixed_cameras(r_arr, t_arr)
    return prob.residuals(None, params_to_points_only(params, n_points))


def cost_func_points_extrinsics(params, n_cameras, n_points, point_3d_indices, camera_indices, k_arr, d_arr, points_2d,
                                project_func=None, device=0):
    params = np.asarray(params, dtype=np.float64)
    prob = SBAProblem(points_2d, point_3d_indices, camera_indices, k_arr, d_arr, n_points, model=camera_model(project_func),
                      device=device)
    return prob.residuals(params[:6 * n_cameras], params[6 * n_cameras:])


def jac_points_extrinsics(params, n_cameras, n_points, point_3d_indices, camera_indices, k_arr, d_arr, points_2d,
                          project_func=None, device=0):
    """Analytic Jacobian of cost_func_points_extrinsics as a scipy.sparse CSR matrix in the reference's
    parameter layout - a drop-in ``jac=`` for scipy.optimize.least_squares."""
    from scipy.sparse import csr_matrix

    params = np.asarray(params, dtype=np.float64)
    prob = SBAProblem(points_2d, point_3d_indices, camera_indices, k_arr, d_arr, n_points, model=camera_model(project_func),
                      device=device)
    _, Jc, Jp, _ = prob.jacobian_blocks(params[:6 * n_cameras], params[6 * n_cameras:])
    n = len(point_3d_indices)
    ci = np.asarray(camera_indices, dtype=np.int64)
    pi = np.asarray(point_3d_indices, dtype=np.int64)
    rows = np.repeat(np.arange(2 * n), 9).reshape(n, 2, 9)
    cols = np.empty((n, 2, 9), dtype=np.int64)
    vals = np.empty((n, 2, 9))
    for k in range(3):
        cols[:, :, k] = (3 * ci + k)[:, None]
        cols[:, :, 3 + k] = (3 * n_cameras + 3 * ci + k)[:, None]
        cols[:, :, 6 + k] = (6 * n_cameras + 3 * pi + k)[:, None]
    vals[:, :, 0:6] = Jc
    vals[:, :, 6:9] = Jp
    return csr_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(2 * n, 6 * n_cameras + 3 * n_points))


def bundle_adjust_points_only(points_2d, points_3d, point_3d_indices, camera_indices, k_arr, d_arr, r_arr, t_arr,
                              project_func=None, f_scale=50, verbose=0, device=None):
    """calib.py:327-341: cauchy f_scale=50, ftol=1e-15, max_nfev=500 -> (obj_pts, residuals)."""
    t0 = time.time()
    out = _solve_sharded(points_2d, points_3d, point_3d_indices, camera_indices, k_arr, d_arr, None, (r_arr, t_arr),
                         f_scale, 500, 1e-15, verbose, model=camera_model(project_func), device=device)
    print("Optimization took {0:.0f} seconds".format(time.time() - t0))
    residuals = dict(before=out["f0"], after=out["fun"])
    return out["pts"], residuals


def bundle_adjust_board_points_only(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr, r_arr, t_arr, triangulate_func=None,
                                    project_func=None, device=None):
    camera_model(project_func, triangulate_func)          # one model for both
    p2, p3, pi, ci = prepare_calib_board_data_for_bundle_adjustment(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr,
                                                                    r_arr, t_arr, triangulate_func)
    return bundle_adjust_points_only(p2, p3, pi, ci, k_arr, d_arr, r_arr, t_arr, project_func, device=device)


def bundle_adjust_points_and_extrinsics(points_2d, points_3d, point_3d_indices, camera_indices, k_arr, d_arr, r_arr, t_arr,
                                        project_func=None, verbose=0, return_info=False, device=None):
    """calib.py:369-390: cauchy f_scale=1, ftol=1e-10, max_nfev=1000
    -> (obj_pts (n,3), r_arr (C,3,3), t_arr (C,3,1), residuals dict(before, after))."""
    n_points = len(points_3d)
    n_cameras = len(k_arr)
    r_vecs = np.array([rodrigues_to_vec(r) for r in r_arr], dtype=np.float64).flatten()
    t_vecs = np.asarray(t_arr, dtype=np.float64).flatten()
    t0 = time.time()
    out = _solve_sharded(points_2d, points_3d, point_3d_indices, camera_indices, k_arr, d_arr,
                         np.concatenate([r_vecs, t_vecs]), None, 1.0, 1000, 1e-10, verbose, model=camera_model(project_func),
                         device=device)
    print("Optimization took {0:.0f} seconds".format(time.time() - t0))
    obj_pts, r_new, t_new = params_to_points_extrinsics(np.concatenate([out["params"], out["pts"].ravel()]), n_cameras,
                                                        n_points)
    residuals = dict(before=out["f0"], after=out["fun"])
    if return_info:
        return obj_pts, r_new, t_new, residuals, out
    return obj_pts, r_new, t_new, residuals


def bundle_adjust_board_points_and_extrinsics(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr, r_arr, t_arr,
                                              triangulate_func=None, project_func=None, device=None):
    camera_model(project_func, triangulate_func)          # one model for both
    p2, p3, pi, ci = prepare_calib_board_data_for_bundle_adjustment(img_pts_arr, fnames_arr, board_shape, k_arr, d_arr,
                                                                    r_arr, t_arr, triangulate_func)
    return bundle_adjust_points_and_extrinsics(p2, p3, pi, ci, k_arr, d_arr, r_arr, t_arr, project_func, device=device)


def _sba_board_points(scene_fpath, points_fpaths, out_fpath, triangulate_func, project_func, device=None):
    """app.py:201-213: load points + scene, refine points and extrinsics, save the ``*_sba.json`` scene."""
    from . import utils

    img_pts_arr, fnames_arr = [], []
    board_shape = None
    for fp in points_fpaths:
        points, fnames, board_shape, _, _ = utils.load_points(fp)
        img_pts_arr.append(points)
        fnames_arr.append(fnames)
    k_arr, d_arr, r_arr, t_arr, cam_res = utils.load_scene(scene_fpath)
    assert len(k_arr) == len(points_fpaths)
    obj_pts, r_new, t_new, res = bundle_adjust_board_points_and_extrinsics(img_pts_arr, fnames_arr, board_shape, k_arr,
                                                                           d_arr, r_arr, t_arr, triangulate_func,
                                                                           project_func, device=device)
    utils.save_scene(out_fpath, k_arr, d_arr, r_new, t_new, cam_res)
    return res


def sba_board_points(scene_fpath, points_fpaths, out_fpath, device=None):
    """app.py:215-218: the standard-camera-model bundle adjustment (cv2.projectPoints / cv2.undistortPoints model)."""
    return _sba_board_points(scene_fpath, points_fpaths, out_fpath, _calib.triangulate_points, _calib.project_points, device)


def sba_board_points_fisheye(scene_fpath, points_fpaths, out_fpath, manual_points_fpath=None,
                             manual_points_only=False, device=None):
    """app.py:220-223 (fisheye model)."""
    from . import utils

    img_pts_arr, fnames_arr = [], []
    board_shape = None
    for fp in points_fpaths:
        points, fnames, board_shape, _, _ = utils.load_points(fp)
        img_pts_arr.append(points)
        fnames_arr.append(fnames)
    k_arr, d_arr, r_arr, t_arr, cam_res = utils.load_scene(scene_fpath)
    assert len(k_arr) == len(img_pts_arr)
    obj_pts, r_new, t_new, res = bundle_adjust_board_points_and_extrinsics(img_pts_arr, fnames_arr, board_shape, k_arr,
                                                                           d_arr, r_arr, t_arr, device=device)
    utils.save_scene(out_fpath, k_arr, d_arr, r_new, t_new, cam_res)
    return res


def sba_points_fisheye(scene_fpath, points_2d_df, device=0):
    """``app.sba_points_fisheye(scene_fpath, points_2d_df) -> (points_3d_df, residuals)`` as called at
    all_optimizations.py:874 and SBA.ipynb:99 (the callee is missing from the reference snapshot; semantics from
    the call sites and from bundle_adjust_points_only, calib.py:327-341): triangulate every (frame, marker) from
    adjacent camera pairs, then refine the 3-D points alone against all cameras that saw them (cameras fixed,
    Cauchy loss, f_scale = 50)."""
    from . import calib, utils

    k_arr, d_arr, r_arr, t_arr, _ = utils.load_scene(scene_fpath)
    assert len(k_arr) == points_2d_df["camera"].nunique()
    points_3d_df = calib.get_pairwise_3d_points_from_df(points_2d_df, k_arr, d_arr, r_arr, t_arr,
                                                        calib.triangulate_points_fisheye, device=device)
    points_3d_df = points_3d_df.reset_index(drop=True)
    points_3d_df["point_index"] = points_3d_df.index
    points_df = points_2d_df.merge(points_3d_df, how="inner", on=["frame", "marker"], suffixes=("_cam", ""))
    points_2d = points_df[["x_cam", "y_cam"]].to_numpy(dtype=np.float32)
    point_indices = points_df["point_index"].to_numpy(dtype=np.int64)
    camera_indices = points_df["camera"].to_numpy(dtype=np.int64)
    points_3d = points_3d_df[["x", "y", "z"]].to_numpy(dtype=np.float32)
    pts, residuals = bundle_adjust_points_only(points_2d, points_3d, point_indices, camera_indices, k_arr, d_arr, r_arr,
                                               t_arr, device=device)
    points_3d_df[["x", "y", "z"]] = pts
    return points_3d_df, residuals
