"""ctypes binding of libacino_b200.so (C ABI declared in include/acino_b200.h).

The shared library is the product; there is no CPU fallback.  Importing this module fails
loudly if the library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``
or ``make``), and creating a handle fails loudly if there is no CUDA device.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacino_b200.so")

N_ACTIVE = 25
N_MARKERS = 20
N_UPPER = 325
MAX_CAMS = 16


class AcinoError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first (make, or __graft_entry__.build()). "
            "acinoset_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.acino_create.argtypes = [ctypes.POINTER(vp), ci]
    lib.acino_create.restype = ci
    lib.acino_destroy.argtypes = [vp]
    lib.acino_destroy.restype = ci
    lib.acino_last_error.argtypes = [vp]
    lib.acino_last_error.restype = ctypes.c_char_p
    lib.acino_version.argtypes = []
    lib.acino_version.restype = ci
    lib.acino_launch_count.argtypes = [vp]
    lib.acino_launch_count.restype = ctypes.c_int64
    lib.acino_set_cameras.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.acino_set_cameras.restype = ci
    lib.acino_set_redescending.argtypes = [vp, cd, cd, cd]
    lib.acino_set_redescending.restype = ci
    lib.acino_fte_eval_dev.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp]
    lib.acino_fte_eval_dev.restype = ci
    lib.acino_fte_eval.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    lib.acino_fte_eval.restype = ci
    lib.acino_fte_eval_kernel_name.argtypes = [ci]
    lib.acino_fte_eval_kernel_name.restype = ctypes.c_char_p
    lib.acino_fk_project_dev.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.acino_fk_project_dev.restype = ci
    lib.acino_fk_project.argtypes = [vp, ci, vp, vp, vp]
    lib.acino_fk_project.restype = ci
    lib.acino_fte_jac_dev.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.acino_fte_jac_dev.restype = ci
    lib.acino_fte_jac.argtypes = [vp, ci, vp, vp, vp]
    lib.acino_fte_jac.restype = ci
    lib.acino_project_points.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    lib.acino_project_points.restype = ci
    lib.acino_undistort_points.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.acino_undistort_points.restype = ci
    lib.acino_triangulate_points.argtypes = [vp, ci] + [vp] * 11
    lib.acino_triangulate_points.restype = ci
    lib.acino_project_points_pinhole.argtypes = [vp, ci, vp, vp, vp, ci, vp, vp, vp]
    lib.acino_project_points_pinhole.restype = ci
    lib.acino_undistort_points_pinhole.argtypes = [vp, ci, vp, vp, vp, ci, ci, vp]
    lib.acino_undistort_points_pinhole.restype = ci
    lib.acino_triangulate_points_pinhole.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp, vp, vp]
    lib.acino_triangulate_points_pinhole.restype = ci
    lib.acino_triangulate_pairwise.argtypes = [vp, ci, ci, vp, vp, vp, vp]
    lib.acino_triangulate_pairwise.restype = ci
    i64 = ctypes.c_int64
    lib.acino_generic_fk.argtypes = [vp, ci, ci, ci] + [vp] * 7
    lib.acino_generic_fk.restype = ci
    u64p = vp
    lib.acino_skel_set.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, u64p, ci, cd, cd, cd, cd]
    lib.acino_skel_set.restype = ci
    lib.acino_skel_eval_dev.argtypes = [vp, ci] + [vp] * 7
    lib.acino_skel_eval_dev.restype = ci
    lib.acino_skel_eval.argtypes = [vp, ci] + [vp] * 6
    lib.acino_skel_eval.restype = ci
    lib.acino_skel_prepare_dev.argtypes = [vp, ci, ci] + [vp] * 9
    lib.acino_skel_prepare_dev.restype = ci
    lib.acino_skel_assemble_dev.argtypes = [vp, ci, vp, vp, vp, vp, cd, vp, vp, vp]
    lib.acino_skel_assemble_dev.restype = ci
    lib.acino_band_solve_dev.argtypes = [vp, i64, ci, vp, vp, vp, vp]
    lib.acino_band_solve_dev.restype = ci
    lib.acino_skel_trial_dev.argtypes = [vp, ci, ci] + [vp] * 6
    lib.acino_skel_trial_dev.restype = ci
    lib.acino_skel_pred_dev.argtypes = [vp, ci] + [vp] * 8
    lib.acino_skel_pred_dev.restype = ci
    lib.acino_stereo_set.argtypes = [vp, ci, ci] + [vp] * 7
    lib.acino_stereo_set.restype = ci
    lib.acino_stereo_set_pinhole.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, ci, vp, vp, ci]
    lib.acino_stereo_set_pinhole.restype = ci
    lib.acino_stereo_init.argtypes = [vp, vp, vp]
    lib.acino_stereo_init.restype = ci
    lib.acino_stereo_step.argtypes = [vp, vp, vp, cd, vp, vp, vp, vp]
    lib.acino_stereo_step.restype = ci
    lib.acino_lm_prepare_dev.argtypes = [vp, ci, i64, i64] + [vp] * 9
    lib.acino_lm_prepare_dev.restype = ci
    lib.acino_lm_assemble_dev.argtypes = [vp, ci, i64, i64, ci, vp, vp, vp, vp, cd, vp, vp, vp, vp]
    lib.acino_lm_assemble_dev.restype = ci
    lib.acino_lm_step_dev.argtypes = [vp, ci, i64, i64] + [vp] * 12
    lib.acino_lm_step_dev.restype = ci
    lib.acino_lm_reduce_dev.argtypes = [vp, ci] + [vp] * 7
    lib.acino_lm_reduce_dev.restype = ci
    lib.acino_bcr_factor_dev.argtypes = [vp, ci] + [vp] * 9
    lib.acino_bcr_factor_dev.restype = ci
    lib.acino_bcr_update_dev.argtypes = [vp, ci] + [vp] * 7
    lib.acino_bcr_update_dev.restype = ci
    lib.acino_bcr_backsub_dev.argtypes = [vp, ci] + [vp] * 7
    lib.acino_bcr_backsub_dev.restype = ci
    lib.acino_sba_cam_bytes.argtypes = []
    lib.acino_sba_cam_bytes.restype = i64
    lib.acino_sba_schur_partial_size.argtypes = [ci, ci]
    lib.acino_sba_schur_partial_size.restype = i64
    lib.acino_sba_cams_dev.argtypes = [vp, ci] + [vp] * 7
    lib.acino_sba_cams_dev.restype = ci
    lib.acino_sba_cams_model_dev.argtypes = [vp, ci, ci, ci] + [vp] * 7
    lib.acino_sba_cams_model_dev.restype = ci
    lib.acino_sba_eval_dev.argtypes = [vp, ci, vp, vp, vp, vp, vp, cd, vp, vp, vp, vp, vp, vp]
    lib.acino_sba_eval_dev.restype = ci
    lib.acino_sba_schur_dev.argtypes = [vp, ci, ci] + [vp] * 7 + [cd, vp, vp, vp, vp]
    lib.acino_sba_schur_dev.restype = ci
    lib.acino_sba_dense_solve_dev.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.acino_sba_dense_solve_dev.restype = ci
    lib.acino_sba_backsub_dev.argtypes = [vp, ci, ci] + [vp] * 7 + [cd, vp, vp, vp, vp, vp]
    lib.acino_sba_backsub_dev.restype = ci
    lib.acino_sba_pred_dev.argtypes = [vp, ci] + [vp] * 10
    lib.acino_sba_pred_dev.restype = ci
    return lib


lib = _load()

# every symbol include/acino_b200.h declares (checked by tests/test_abi.py without a GPU)
EXPORTED = [
    "acino_create", "acino_destroy", "acino_last_error", "acino_version", "acino_launch_count",
    "acino_set_cameras", "acino_set_redescending", "acino_fte_eval_dev", "acino_fte_eval", "acino_fte_eval_kernel_name",
    "acino_fk_project_dev", "acino_fk_project", "acino_fte_jac_dev", "acino_fte_jac", "acino_project_points", "acino_undistort_points",
    "acino_triangulate_points", "acino_triangulate_pairwise", "acino_generic_fk",
    "acino_project_points_pinhole", "acino_undistort_points_pinhole", "acino_triangulate_points_pinhole",
    "acino_skel_set", "acino_skel_eval_dev", "acino_skel_eval", "acino_skel_prepare_dev", "acino_skel_assemble_dev",
    "acino_band_solve_dev", "acino_skel_trial_dev", "acino_skel_pred_dev",
    "acino_stereo_set", "acino_stereo_set_pinhole", "acino_stereo_init", "acino_stereo_step",
    "acino_lm_prepare_dev", "acino_lm_assemble_dev", "acino_lm_step_dev", "acino_lm_reduce_dev",
    "acino_lm_desc_size", "acino_lm_plan_create", "acino_lm_plan_destroy", "acino_lm_enqueue",
    "acino_bcr_factor_dev", "acino_bcr_update_dev", "acino_bcr_backsub_dev",
    "acino_sba_cam_bytes", "acino_sba_schur_partial_size", "acino_sba_cams_dev", "acino_sba_cams_model_dev", "acino_sba_eval_dev",
    "acino_sba_schur_dev", "acino_sba_dense_solve_dev", "acino_sba_backsub_dev", "acino_sba_pred_dev",
]


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _host(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class Handle:
    """One handle per GPU: owns the camera table, loss constants and a staging workspace."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib.acino_create(ctypes.byref(self._h), int(device))
        if rc != 0:
            raise AcinoError(f"acino_create failed ({rc}): {lib.acino_last_error(None).decode()}")
        self.device = int(device)
        self.n_cams = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.acino_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise AcinoError(f"{what} failed ({rc}): {lib.acino_last_error(self._h).decode()}")

    @property
    def launch_count(self):
        return int(lib.acino_launch_count(self._h))

    @staticmethod
    def fte_kernel_name(n_frames=1 << 20):
        return lib.acino_fte_eval_kernel_name(int(n_frames)).decode()

    # ---- scene
    def set_cameras(self, K, D, R, t):
        K = _host(K, np.float64)
        C = K.shape[0]
        K = K.reshape(C, 9)
        D = _host(D, np.float64).reshape(C, 4)
        R = _host(R, np.float64).reshape(C, 9)
        t = _host(t, np.float64).reshape(C, 3)
        self._check(lib.acino_set_cameras(self._h, C, _np_ptr(K), _np_ptr(D), _np_ptr(R), _np_ptr(t)),
                    "acino_set_cameras")
        self.n_cams = C
        # same table again (every TRI / SBA call re-installs its scene): not a change; solvers that live across calls
        # (lm.FTESolver, ekf) remember the version they were built for and refuse to run against another scene
        sig = (K.tobytes(), D.tobytes(), R.tobytes(), t.tobytes())
        if sig != getattr(self, "_scene_sig", None):
            self._scene_sig = sig
            self.scene_version = getattr(self, "scene_version", 0) + 1

    def check_scene(self, version, who):
        if version != getattr(self, "scene_version", 0):
            raise AcinoError(f"{who}: the camera table of the handle on cuda:{self.device} was replaced (scene version "
                             f"{getattr(self, 'scene_version', 0)}, expected {version}) - a handle holds ONE scene; use one "
                             "handle per scene (acinoset_b200.Handle(device)) or re-create the solver")

    def set_redescending(self, a, b, c):
        self._check(lib.acino_set_redescending(self._h, float(a), float(b), float(c)), "acino_set_redescending")

    # ---- host-buffer API (copies inside the call)
    def fte_eval(self, x, meas, w, want_H=True, out=None):
        x = _host(x, np.float32)
        N = x.shape[0]
        C = self.n_cams
        x = _host(x, np.float32, (N, N_ACTIVE))
        meas = _host(meas, np.float32, (N, C, N_MARKERS, 2))
        w = _host(w, np.float32, (N, C, N_MARKERS))
        if out is None:
            cost = np.empty(N, np.float32)
            g = np.empty((N, N_ACTIVE), np.float32)
            H = np.empty((N, N_UPPER), np.float32) if want_H else None
        else:
            cost, g, H = out
        self._check(lib.acino_fte_eval(self._h, N, _np_ptr(x), _np_ptr(meas), _np_ptr(w), _np_ptr(cost),
                                       _np_ptr(g), _np_ptr(H)), "acino_fte_eval")
        return cost, g, H

    def fk_project(self, x, want_pos=True, want_uv=True):
        x = _host(x, np.float32)
        N = x.shape[0]
        x = _host(x, np.float32, (N, N_ACTIVE))
        pos = np.empty((N, N_MARKERS, 3), np.float32) if want_pos else None
        uv = np.empty((N, self.n_cams, N_MARKERS, 2), np.float32) if want_uv else None
        self._check(lib.acino_fk_project(self._h, N, _np_ptr(x), _np_ptr(pos), _np_ptr(uv)), "acino_fk_project")
        return pos, uv

    def fte_jac(self, x, want_uv=True, want_J=True):
        """h(x) and d h / d x for every camera: x (N,25) -> uv (N,C,20,2), J (N,C,20,2,25)."""
        x = _host(x, np.float32)
        N = x.shape[0]
        x = _host(x, np.float32, (N, N_ACTIVE))
        uv = np.empty((N, self.n_cams, N_MARKERS, 2), np.float32) if want_uv else None
        J = np.empty((N, self.n_cams, N_MARKERS, 2, N_ACTIVE), np.float32) if want_J else None
        self._check(lib.acino_fte_jac(self._h, N, _np_ptr(x), _np_ptr(uv), _np_ptr(J)), "acino_fte_jac")
        return uv, J

    # ---- camera geometry (fp64, host buffers)
    @staticmethod
    def _cam(K, D, R=None, t=None):
        K = _host(K, np.float64).reshape(9)
        D = _host(D, np.float64).reshape(4)
        R = None if R is None else _host(R, np.float64).reshape(9)
        t = None if t is None else _host(t, np.float64).reshape(3)
        return K, D, R, t

    def project_points(self, X, K, D, R, t):
        X = _host(X, np.float64).reshape(-1, 3)
        K, D, R, t = self._cam(K, D, R, t)
        uv = np.empty((X.shape[0], 2), np.float64)
        self._check(lib.acino_project_points(self._h, X.shape[0], _np_ptr(X), _np_ptr(K), _np_ptr(D), _np_ptr(R),
                                             _np_ptr(t), _np_ptr(uv)), "acino_project_points")
        return uv

    def undistort_points(self, uv, K, D):
        uv = _host(uv, np.float64).reshape(-1, 2)
        K, D, _, _ = self._cam(K, D)
        out = np.empty_like(uv)
        self._check(lib.acino_undistort_points(self._h, uv.shape[0], _np_ptr(uv), _np_ptr(K), _np_ptr(D), _np_ptr(out)),
                    "acino_undistort_points")
        return out

    def triangulate_points(self, uv1, uv2, cam1, cam2):
        uv1 = _host(uv1, np.float64).reshape(-1, 2)
        uv2 = _host(uv2, np.float64).reshape(-1, 2)
        if uv1.shape != uv2.shape:
            raise ValueError("img_pts_1 and img_pts_2 must hold the same number of points")
        K1, D1, R1, t1 = self._cam(*cam1)
        K2, D2, R2, t2 = self._cam(*cam2)
        X = np.empty((uv1.shape[0], 3), np.float64)
        self._check(lib.acino_triangulate_points(self._h, uv1.shape[0], _np_ptr(uv1), _np_ptr(uv2), _np_ptr(K1),
                                                 _np_ptr(D1), _np_ptr(R1), _np_ptr(t1), _np_ptr(K2), _np_ptr(D2),
                                                 _np_ptr(R2), _np_ptr(t2), _np_ptr(X)), "acino_triangulate_points")
        return X

    # ---- generic-skeleton FTE (build.py variant), fp64
    def skel_set(self, flat, loss="abs", abc=(3.0, 10.0, 20.0), delta=0.05):
        """flat: acinoset_b200.skeleton.flatten_skeleton(skel_dict); cameras must be set first."""
        n_parts, n_links, n_out = len(flat["parts"]), len(flat["link_parent"]), len(flat["out_order"])
        kind = {"redescending": 0, "abs": 1}[loss]
        dm = _host(flat["dof_mask"], np.int32)
        lp, lf = _host(flat["link_parent"], np.int32), _host(flat["link_flags"], np.int32)
        tv, path = _host(flat["link_tv"], np.float64), _host(flat["out_path"], np.uint64)
        self._check(lib.acino_skel_set(self._h, n_parts, n_links, n_out, _np_ptr(dm), _np_ptr(lp), _np_ptr(lf), _np_ptr(tv),
                                       _np_ptr(path), kind, float(abc[0]), float(abc[1]), float(abc[2]), float(delta)),
                    "acino_skel_set")
        self.skel_shape = (3 + 3 * n_parts, n_out)

    def skel_eval(self, x, meas, w, want_H=True):
        """x (N,P), meas (N,C,n_out,2), w (N,C,n_out) -> cost (N,), g (N,P), H (N,P(P+1)/2) fp64 (host buffers)."""
        P, n_out = self.skel_shape
        x = _host(x, np.float64)
        N = x.shape[0]
        x = _host(x, np.float64, (N, P))
        meas = _host(meas, np.float64, (N, self.n_cams, n_out, 2))
        w = _host(w, np.float64, (N, self.n_cams, n_out))
        cost = np.empty(N)
        g = np.empty((N, P))
        H = np.empty((N, P * (P + 1) // 2)) if want_H else None
        self._check(lib.acino_skel_eval(self._h, N, _np_ptr(x), _np_ptr(meas), _np_ptr(w), _np_ptr(cost), _np_ptr(g),
                                        _np_ptr(H)), "acino_skel_eval")
        return cost, g, H

    # ---- pinhole twins (cv2.projectPoints / cv2.undistortPoints / triangulate_points)
    @staticmethod
    def _pincam(K, d, R=None, t=None):
        K = _host(K, np.float64).reshape(9)
        d = np.zeros(0) if d is None else _host(d, np.float64).reshape(-1)
        R = None if R is None else _host(R, np.float64).reshape(9)
        t = None if t is None else _host(t, np.float64).reshape(3)
        return K, d, R, t

    def project_points_pinhole(self, X, K, d, R, t):
        X = _host(X, np.float64).reshape(-1, 3)
        K, d, R, t = self._pincam(K, d, R, t)
        uv = np.empty((X.shape[0], 2), np.float64)
        self._check(lib.acino_project_points_pinhole(self._h, X.shape[0], _np_ptr(X), _np_ptr(K), _np_ptr(d), d.size,
                                                     _np_ptr(R), _np_ptr(t), _np_ptr(uv)), "acino_project_points_pinhole")
        return uv

    def undistort_points_pinhole(self, uv, K, d, to_pixels=False):
        uv = _host(uv, np.float64).reshape(-1, 2)
        K, d, _, _ = self._pincam(K, d)
        out = np.empty_like(uv)
        self._check(lib.acino_undistort_points_pinhole(self._h, uv.shape[0], _np_ptr(uv), _np_ptr(K), _np_ptr(d), d.size,
                                                       1 if to_pixels else 0, _np_ptr(out)), "acino_undistort_points_pinhole")
        return out

    def triangulate_points_pinhole(self, uv1, uv2, cam1, cam2):
        uv1 = _host(uv1, np.float64).reshape(-1, 2)
        uv2 = _host(uv2, np.float64).reshape(-1, 2)
        if uv1.shape != uv2.shape:
            raise ValueError("img_pts_1 and img_pts_2 must hold the same number of points")
        K1, d1, R1, t1 = self._pincam(*cam1)
        K2, d2, R2, t2 = self._pincam(*cam2)
        X = np.empty((uv1.shape[0], 3), np.float64)
        self._check(lib.acino_triangulate_points_pinhole(self._h, uv1.shape[0], _np_ptr(uv1), _np_ptr(uv2), _np_ptr(K1),
                                                         _np_ptr(d1), d1.size, _np_ptr(R1), _np_ptr(t1), _np_ptr(K2),
                                                         _np_ptr(d2), d2.size, _np_ptr(R2), _np_ptr(t2), _np_ptr(X)),
                    "acino_triangulate_points_pinhole")
        return X

    def triangulate_pairwise(self, uv, valid):
        uv = _host(uv, np.float64)
        N, C, L, _ = uv.shape
        if C != self.n_cams:
            raise ValueError(f"uv has {C} cameras, the handle has {self.n_cams}")
        valid = _host(valid, np.uint8, (N, C, L))
        pos = np.empty((N, L, 3), np.float64)
        cnt = np.empty((N, L), np.int32)
        self._check(lib.acino_triangulate_pairwise(self._h, N, L, _np_ptr(uv), _np_ptr(valid), _np_ptr(pos), _np_ptr(cnt)),
                    "acino_triangulate_pairwise")
        return pos, cnt

    # ---- device-pointer API (torch tensors on this handle's device; stream-ordered)
    def fte_eval_dev(self, x, meas, w, cost=None, g=None, H=None, stream=None):
        import torch

        N = x.shape[0]
        for tns, shp in ((x, (N, N_ACTIVE)), (meas, (N, self.n_cams, N_MARKERS, 2)), (w, (N, self.n_cams, N_MARKERS))):
            _check_dev(tns, shp, self.device)
        for tns, shp in ((cost, (N,)), (g, (N, N_ACTIVE)), (H, (N, N_UPPER))):
            if tns is not None:
                _check_dev(tns, shp, self.device)
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(lib.acino_fte_eval_dev(self._h, N, _dp(x), _dp(meas), _dp(w), _dp(cost), _dp(g), _dp(H),
                                           ctypes.c_void_p(s)), "acino_fte_eval_dev")

    def fk_project_dev(self, x, pos=None, uv=None, stream=None):
        import torch

        N = x.shape[0]
        _check_dev(x, (N, N_ACTIVE), self.device)
        if pos is not None:
            _check_dev(pos, (N, N_MARKERS, 3), self.device)
        if uv is not None:
            _check_dev(uv, (N, self.n_cams, N_MARKERS, 2), self.device)
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(lib.acino_fk_project_dev(self._h, N, _dp(x), _dp(pos), _dp(uv), ctypes.c_void_p(s)),
                    "acino_fk_project_dev")

    def fte_jac_dev(self, x, uv=None, J=None, stream=None):
        import torch

        N = x.shape[0]
        _check_dev(x, (N, N_ACTIVE), self.device)
        if uv is not None:
            _check_dev(uv, (N, self.n_cams, N_MARKERS, 2), self.device)
        if J is not None:
            _check_dev(J, (N, self.n_cams, N_MARKERS, 2, N_ACTIVE), self.device)
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(lib.acino_fte_jac_dev(self._h, N, _dp(x), _dp(uv), _dp(J), ctypes.c_void_p(s)), "acino_fte_jac_dev")

    # ---- raw device-pointer calls used by acinoset_b200.lm (arguments are torch tensors / scalars)
    def call_dev(self, name, *args, stream=None):
        import torch

        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        conv = []
        for a in args:
            if a is None:
                conv.append(None)
            elif hasattr(a, "data_ptr"):
                conv.append(ctypes.c_void_p(a.data_ptr()))
            else:
                conv.append(a)
        self._check(getattr(lib, name)(self._h, *conv, ctypes.c_void_p(s)), name)


def _dp(tns):
    return ctypes.c_void_p(tns.data_ptr()) if tns is not None else None


def _check_dev(tns, shape, device):
    import torch

    if not tns.is_cuda or tns.device.index != device:
        raise ValueError(f"tensor must live on cuda:{device}")
    if tns.dtype != torch.float32 or not tns.is_contiguous() or tuple(tns.shape) != tuple(shape):
        raise ValueError(f"expected contiguous float32 tensor of shape {shape}, got {tns.dtype} {tuple(tns.shape)}")
