"""Generic skeleton-pickle forward kinematics - the `pose_to_3d` of the reference's data-driven
builder (/root/reference/src/build.py:19-95) on the GPU.

`build_pose_function(skel_dict)` simulates the builder's dictionary manipulations ONCE on the host
(which part is a root, which local rotation each link uses and whether it is transposed at that
point - the reference toggles ``rot_dict[child + "_i"]`` each time a part appears as a child,
build.py:78-79 - and the pose_dict insertion order of the output rows) and returns a function that
evaluates any number of frames with one kernel launch (`acino_generic_fk`).
"""
import ctypes
import pickle

import numpy as np

from . import fte as _fte
from . import _lib


def load_skeleton(skel_file):
    """build.py:19-26"""
    with open(skel_file, "rb") as handle:
        return pickle.load(handle)


def flatten_skeleton(skel_dict):
    """-> dict(parts, dof_mask, link_parent, link_child, link_flags, link_tv, out_order).

    out_order: part indices in pose_dict insertion order (the row order of the reference's
    ``t_poses_mat``, build.py:82-86)."""
    links = skel_dict["links"]
    positions = skel_dict["positions"]
    dofs = {k: list(v) for k, v in skel_dict["dofs"].items()}
    for joint in skel_dict["markers"]:          # build.py:37-38
        dofs[joint] = [1, 1, 1]
    parts = list(dofs.keys())                   # angle index i = position in the dofs dict (build.py:51-62)
    idx = {p: i for i, p in enumerate(parts)}
    if len(parts) > 32:
        raise ValueError("generic FK supports up to 32 parts")
    if len(parts) != len(positions):
        raise ValueError("dofs and positions must describe the same parts")
    dof_mask = np.array([(1 if dofs[p][0] else 0) | (2 if dofs[p][1] else 0) | (4 if dofs[p][2] else 0) for p in parts],
                        dtype=np.int32)
    transposed = {p: True for p in parts}       # rot_dict[p + "_i"] starts as the transpose (build.py:61)
    out_order, lp, lc, lf, tv = [], [], [], [], []
    path = {}                                   # part -> bitmask of the links whose increments sum to its pose
    for link in links:
        if len(link) == 1:
            if link[0] not in out_order:
                out_order.append(link[0])
            path[link[0]] = 0
            continue
        a, b = link
        if a not in out_order:
            out_order.append(a)
            path[a] = 0
        transposed[b] = not transposed[b]       # build.py:79 (before the pose of b is formed)
        path[b] = path[a] | (1 << len(lp))      # build.py:80: pose[b] = pose[a] (as it is NOW) + this link's increment
        lp.append(idx[a])
        lc.append(idx[b])
        lf.append(1 if transposed[a] else 0)
        tv.append([positions[b][k] - positions[a][k] for k in range(3)])
        if b not in out_order:
            out_order.append(b)
    return dict(parts=parts, dof_mask=dof_mask, link_parent=np.array(lp, dtype=np.int32),
                link_child=np.array(lc, dtype=np.int32), link_flags=np.array(lf, dtype=np.int32),
                link_tv=np.array(tv, dtype=np.float64).reshape(-1, 3), out_order=[idx[p] for p in out_order],
                out_names=out_order, out_path=np.array([path[p] for p in out_order], dtype=np.uint64))


def build_pose_function(skel_dict, device=0):
    """Returns pose_to_3d(x): x (3 + 3L,) or (N, 3 + 3L) -> (n_out, 3) or (N, n_out, 3) float64,
    rows in the reference's pose_dict order (``.out_names`` on the returned function)."""
    flat = flatten_skeleton(skel_dict)
    P = len(flat["parts"])
    nl = len(flat["link_parent"])
    order = np.array(flat["out_order"], dtype=np.int64)

    def p(a):
        return ctypes.c_void_p(a.ctypes.data) if a.size else None

    def pose_to_3d(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        single = x.ndim == 1
        x2 = np.atleast_2d(x)
        if x2.shape[1] != 3 + 3 * P:
            raise ValueError(f"state must have {3 + 3 * P} entries [x,y,z,*phi,*theta,*psi]")
        h = _fte.get_handle(device)
        pos = np.empty((x2.shape[0], P, 3), dtype=np.float64)
        h._check(_lib.lib.acino_generic_fk(h._h, x2.shape[0], P, nl, p(flat["dof_mask"]), p(flat["link_parent"]),
                                           p(flat["link_child"]), p(flat["link_flags"]), p(flat["link_tv"]), p(x2), p(pos)),
                 "acino_generic_fk")
        out = pos[:, order]
        return out[0] if single else out

    pose_to_3d.out_names = flat["out_names"]
    return pose_to_3d
