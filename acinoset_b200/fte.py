"""Full Trajectory Estimation (FTE) on the GPU - host side.

Mirrors the FTE section of the reference (src/all_optimizations.py:22-566): the same
measurement model (pt3d_to_2d :193-209), weights (:302-308), redescending loss (:25-27,497),
dynamics/smoothness term (:369-391, :245-252) and pose bounds (:403-483); the residual /
Jacobian evaluation and the solve run in libacino_b200.so.
"""
import numpy as np

from . import _lib

N_ACTIVE = _lib.N_ACTIVE
N_MARKERS = _lib.N_MARKERS
N_UPPER = _lib.N_UPPER

# marker order = rows of `positions` (all_optimizations.py:170-178)
MARKERS = [
    "l_eye", "r_eye", "nose", "neck_base", "spine", "tail_base", "tail1", "tail2",
    "l_shoulder", "l_front_knee", "l_front_ankle", "r_shoulder", "r_front_knee",
    "r_front_ankle", "l_hip", "l_back_knee", "l_back_ankle", "r_hip", "r_back_knee",
    "r_back_ankle",
]
# indices of the 25 active slots in the reference's 45-vector [x,y,z,*phi(14),*theta(14),*psi(14)]
ACTIVE_IDX = np.array([0, 1, 2, 3, 4, 6] + list(range(17, 31)) + [31, 32, 34, 35, 36])
MEAS_SIGMA_R = 5.0  # all_optimizations.py:243

_handles = {}


def get_handle(device=0):
    """One cached library handle per GPU."""
    h = _handles.get(device)
    if h is None:
        h = _lib.Handle(device)
        _handles[device] = h
    return h


def set_scene(k_arr, d_arr, r_arr, t_arr, device=0):
    """Install the camera table (the arrays utils.load_scene / find_scene_file return)."""
    h = get_handle(device)
    h.set_cameras(k_arr, np.asarray(d_arr).reshape(-1, 4), r_arr, np.asarray(t_arr).reshape(-1, 3))
    return h


def meas_weights(likelihood, dlc_thresh, sigma=MEAS_SIGMA_R):
    """init_meas_weights (all_optimizations.py:302-308): 1/R if likelihood > thresh else 0."""
    return np.where(np.asarray(likelihood) > dlc_thresh, np.float32(1.0 / sigma), np.float32(0.0)).astype(np.float32)


def fte_eval(x, meas, w, device=0, want_H=True):
    """Residual + Jacobian evaluation of N frames through the C ABI with HOST buffers.

    x (N,25), meas (N,C,20,2), w (N,C,20) -> cost (N,), g (N,25), H (N,325) packed upper."""
    return get_handle(device).fte_eval(x, meas, w, want_H=want_H)


def pose_to_3d(x, device=0):
    """pose_to_3d (all_optimizations.py:186) for a batch of active states: (N,25) -> (N,20,3)."""
    x = np.asarray(x, dtype=np.float32)
    single = x.ndim == 1
    pos, _ = get_handle(device).fk_project(np.atleast_2d(x), want_pos=True, want_uv=False)
    return pos[0] if single else pos


def unpack_upper(Hu):
    Hu = np.asarray(Hu)
    iu = np.triu_indices(N_ACTIVE)
    H = np.zeros(Hu.shape[:-1] + (N_ACTIVE, N_ACTIVE), dtype=Hu.dtype)
    H[..., iu[0], iu[1]] = Hu
    H[..., iu[1], iu[0]] = Hu
    return H
