"""Oracle (test infrastructure): cheetah forward kinematics + analytic Jacobian, and the
generic skeleton-pickle FK builder, in fp64 NumPy.

Follows
  * rot_x / rot_y / rot_z (passive rotations)     /root/reference/src/all_optimizations.py:66-91
  * rotation chain RI_0 .. RI_13                   all_optimizations.py:101-128
  * marker positions p_* and their row order       all_optimizations.py:138-179
  * state order sym_list = [x,y,z,*phi,*theta,*psi] all_optimizations.py:182-185
  * generic builder (with its quirks)              /root/reference/src/build.py:32-95
"""
import numpy as np

N_JOINTS = 14
N_STATE_FULL = 3 + 3 * N_JOINTS  # 45 (all_optimizations.py:288)
# indices (0-based, in the 45-vector) of the state slots the FK actually depends on:
# x,y,z, phi0,phi1,phi3, theta0..13, psi0,psi1,psi3,psi4,psi5.  Identical to the
# non-zero entries of Q (all_optimizations.py:245-250) and to the column order of the
# saved result pickle (all_optimizations.py:540-556).
ACTIVE_IDX = np.array([0, 1, 2, 3, 4, 6] + list(range(17, 31)) + [31, 32, 34, 35, 36])
N_ACTIVE = 25

MARKERS = [
    "l_eye", "r_eye", "nose", "neck_base", "spine", "tail_base", "tail1", "tail2",
    "l_shoulder", "l_front_knee", "l_front_ankle", "r_shoulder", "r_front_knee",
    "r_front_ankle", "l_hip", "l_back_knee", "l_back_ankle", "r_hip", "r_back_knee",
    "r_back_ankle",
]
N_MARKERS = 20

# measurement model std-devs, all_optimizations.py:245-252 (Q = sigma**2; weight 1/Q, 0 where sigma==0)
Q_SIGMA = np.array(
    [4, 7, 5]
    + [13, 32, 0, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    + [9, 18, 43, 53, 90, 118, 247, 186, 194, 164, 295, 243, 334, 149]
    + [26, 12, 0, 34, 43, 51, 0, 0, 0, 0, 0, 0, 0, 0],
    dtype=np.float64,
)

# box bounds of all_optimizations.py:403-483, as (0-based index into the 45-vector, lo, hi)
_P6, _P15, _P2 = np.pi / 6, np.pi / 1.5, np.pi / 2
BOUNDS_FULL = [
    (3, -_P6, _P6), (17, -_P6, _P6),                      # head phi0, theta0
    (4, -_P6, _P6), (18, -_P6, _P6), (32, -_P6, _P6),     # neck phi1, theta1, psi1
    (19, -_P6, _P6),                                      # front torso theta2
    (20, -_P6, _P6), (6, -_P6, _P6), (34, -_P6, _P6),     # back torso theta3, phi3, psi3
    (21, -_P15, _P15), (35, -_P15, _P15),                 # tail base theta4, psi4
    (22, -_P15, _P15), (36, -_P15, _P15),                 # tail mid theta5, psi5
    (23, -_P2, _P2), (24, -np.pi, 0.0),                   # l_shoulder theta6, l_front_knee theta7
    (25, -_P2, _P2), (26, -np.pi, 0.0),                   # r_shoulder theta8, r_front_knee theta9
    (27, -_P2, _P2), (28, 0.0, np.pi),                    # l_hip theta10, l_back_knee theta11
    (29, -_P2, _P2), (30, 0.0, np.pi),                    # r_hip theta12, r_back_knee theta13
]


def active_bounds():
    """(lo[25], hi[25]) in the active ordering; +-inf where unbounded."""
    lo = np.full(N_ACTIVE, -np.inf)
    hi = np.full(N_ACTIVE, np.inf)
    pos = {int(f): i for i, f in enumerate(ACTIVE_IDX)}
    for f, l, h in BOUNDS_FULL:
        lo[pos[f]] = l
        hi[pos[f]] = h
    return lo, hi


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    o, z = np.ones_like(c), np.zeros_like(c)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, s], -1), np.stack([z, -s, c], -1)], -2)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    o, z = np.ones_like(c), np.zeros_like(c)
    return np.stack([np.stack([c, z, -s], -1), np.stack([z, o, z], -1), np.stack([s, z, c], -1)], -2)


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    o, z = np.ones_like(c), np.zeros_like(c)
    return np.stack([np.stack([c, s, z], -1), np.stack([-s, c, z], -1), np.stack([z, z, o], -1)], -2)


def full_from_active(xa):
    xa = np.asarray(xa, dtype=np.float64)
    x = np.zeros(xa.shape[:-1] + (N_STATE_FULL,))
    x[..., ACTIVE_IDX] = xa
    return x


def cheetah_rotations(x45):
    """RI_k (inertial -> body k), k = 0..13, for a batch of full 45-states (...,45)."""
    x45 = np.asarray(x45, dtype=np.float64)
    phi = x45[..., 3:17]
    th = x45[..., 17:31]
    psi = x45[..., 31:45]
    RI = [None] * N_JOINTS
    RI[0] = rot_z(psi[..., 0]) @ rot_x(phi[..., 0]) @ rot_y(th[..., 0])
    RI[1] = rot_z(psi[..., 1]) @ rot_x(phi[..., 1]) @ rot_y(th[..., 1]) @ RI[0]
    RI[2] = rot_y(th[..., 2]) @ RI[1]
    RI[3] = rot_z(psi[..., 3]) @ rot_x(phi[..., 3]) @ rot_y(th[..., 3]) @ RI[2]
    RI[4] = rot_z(psi[..., 4]) @ rot_y(th[..., 4]) @ RI[3]
    RI[5] = rot_z(psi[..., 5]) @ rot_y(th[..., 5]) @ RI[4]
    RI[6] = rot_y(th[..., 6]) @ RI[2]
    RI[7] = rot_y(th[..., 7]) @ RI[6]
    RI[8] = rot_y(th[..., 8]) @ RI[2]
    RI[9] = rot_y(th[..., 9]) @ RI[8]
    RI[10] = rot_y(th[..., 10]) @ RI[3]
    RI[11] = rot_y(th[..., 11]) @ RI[10]
    RI[12] = rot_y(th[..., 12]) @ RI[3]
    RI[13] = rot_y(th[..., 13]) @ RI[12]
    return RI


# (marker, parent marker or None for the head point, joint whose R_k_I rotates the offset, offset)
SEGMENTS = [
    ("l_eye", None, 0, (0.0, 0.03, 0.0)),
    ("r_eye", None, 0, (0.0, -0.03, 0.0)),
    ("nose", None, 0, (0.055, 0.0, -0.055)),
    ("neck_base", None, 1, (-0.28, 0.0, 0.0)),
    ("spine", "neck_base", 2, (-0.37, 0.0, 0.0)),
    ("tail_base", "spine", 3, (-0.37, 0.0, 0.0)),
    ("tail1", "tail_base", 4, (-0.28, 0.0, 0.0)),
    ("tail2", "tail1", 5, (-0.36, 0.0, 0.0)),
    ("l_shoulder", "neck_base", 2, (-0.04, 0.08, -0.10)),
    ("l_front_knee", "l_shoulder", 6, (0.0, 0.0, -0.24)),
    ("l_front_ankle", "l_front_knee", 7, (0.0, 0.0, -0.28)),
    ("r_shoulder", "neck_base", 2, (-0.04, -0.08, -0.10)),
    ("r_front_knee", "r_shoulder", 8, (0.0, 0.0, -0.24)),
    ("r_front_ankle", "r_front_knee", 9, (0.0, 0.0, -0.28)),
    ("l_hip", "tail_base", 3, (0.12, 0.08, -0.06)),
    ("l_back_knee", "l_hip", 10, (0.0, 0.0, -0.32)),
    ("l_back_ankle", "l_back_knee", 11, (0.0, 0.0, -0.25)),
    ("r_hip", "tail_base", 3, (0.12, -0.08, -0.06)),
    ("r_back_knee", "r_hip", 12, (0.0, 0.0, -0.32)),
    ("r_back_ankle", "r_back_knee", 13, (0.0, 0.0, -0.25)),
]


def cheetah_fk(x45):
    """pose_to_3d: (...,45) -> (...,20,3) marker positions in the inertial frame."""
    x45 = np.asarray(x45, dtype=np.float64)
    RI = cheetah_rotations(x45)
    head = x45[..., 0:3]
    pos = {}
    for name, parent, k, off in SEGMENTS:
        base = head if parent is None else pos[parent]
        RkI = np.swapaxes(RI[k], -1, -2)  # R_k_I = RI_k.T
        pos[name] = base + RkI @ np.asarray(off, dtype=np.float64)
    return np.stack([pos[m] for m in MARKERS], axis=-2)


def cheetah_fk_active(xa):
    return cheetah_fk(full_from_active(xa))


def cheetah_fk_jac_fd(xa, eps=1e-6):
    """Central-difference Jacobian d positions / d active state: (...,20,3,25)."""
    xa = np.asarray(xa, dtype=np.float64)
    J = np.zeros(xa.shape[:-1] + (N_MARKERS, 3, N_ACTIVE))
    for p in range(N_ACTIVE):
        d = np.zeros(N_ACTIVE)
        d[p] = eps
        J[..., p] = (cheetah_fk_active(xa + d) - cheetah_fk_active(xa - d)) / (2 * eps)
    return J


# ---- analytic Jacobian -------------------------------------------------------------
# joint parents in the rotation chain (all_optimizations.py:101-128)
JOINT_PARENT = [-1, 0, 1, 2, 3, 4, 2, 6, 2, 8, 3, 10, 3, 12]
# joint whose rotation moves each marker's last segment
MARKER_JOINT = [s[2] for s in SEGMENTS]
# pivot of a joint = the point the first segment rotated by that joint starts from:
# None -> head point, else a marker name
JOINT_PIVOT = [None, None, "neck_base", "spine", "tail_base", "tail1", "l_shoulder",
               "l_front_knee", "r_shoulder", "r_front_knee", "l_hip", "l_back_knee",
               "r_hip", "r_back_knee"]


def _joint_ancestors(k):
    out = []
    while k >= 0:
        out.append(k)
        k = JOINT_PARENT[k]
    return out


def active_slot_info():
    """For each of the 25 active slots: (kind, joint) with kind in 'x','y','z','phi','theta','psi'."""
    info = []
    for f in ACTIVE_IDX:
        f = int(f)
        if f < 3:
            info.append(("xyz"[f], -1))
        elif f < 17:
            info.append(("phi", f - 3))
        elif f < 31:
            info.append(("theta", f - 17))
        else:
            info.append(("psi", f - 31))
    return info


def cheetah_fk_jac(xa):
    """Analytic d positions / d active state, (...,20,3,25).

    With R_k_I = R_parent_I * Ry_a(theta) Rx_a(phi) Rz_a(psi) (active rotations; the
    reference's passive matrices are their transposes), d(R_k_I v)/d alpha =
    omega_alpha x (R_k_I v) where the world-frame axes are
        omega_theta = R_parent_I e_y,  omega_phi = R_parent_I Ry_a(theta) e_x,
        omega_psi = R_k_I e_z,
    so d p_l / d alpha = omega_alpha x (p_l - pivot(joint(alpha))) for every marker l
    downstream of the joint, 0 otherwise.
    """
    xa = np.asarray(xa, dtype=np.float64)
    x45 = full_from_active(xa)
    RI = cheetah_rotations(x45)
    P = cheetah_fk(x45)
    head = x45[..., 0:3]
    th = x45[..., 17:31]
    J = np.zeros(xa.shape[:-1] + (N_MARKERS, 3, N_ACTIVE))
    eye = np.eye(3)
    for p, (kind, k) in enumerate(active_slot_info()):
        if k < 0:
            J[..., :, :, p] = eye["xyz".index(kind)]
            continue
        par = JOINT_PARENT[k]
        Rpar_I = np.swapaxes(RI[par], -1, -2) if par >= 0 else np.broadcast_to(eye, RI[0].shape)
        if kind == "theta":
            om = Rpar_I[..., :, 1]
        elif kind == "phi":
            # Ry_a(theta) e_x = first column of rot_y(theta).T = first row of the passive rot_y
            om = (Rpar_I @ rot_y(th[..., k])[..., 0, :, None])[..., 0]
        else:
            om = np.swapaxes(RI[k], -1, -2)[..., :, 2]
        piv = head if JOINT_PIVOT[k] is None else P[..., MARKERS.index(JOINT_PIVOT[k]), :]
        for l in range(N_MARKERS):
            if k in _joint_ancestors(MARKER_JOINT[l]):
                J[..., l, :, p] = np.cross(om, P[..., l, :] - piv)
    return J


# ---- generic skeleton builder (build.py:32-95), quirks preserved --------------------
def generic_fk_builder(skel_dict):
    """Return pose_to_3d(x) for a skeleton pickle dict, x = [x,y,z,*phi(L),*theta(L),*psi(L)].

    Replicates build.py:32-95 step by step, including:
      * dofs of every name in ``markers`` forced to [1,1,1]                (build.py:37-38)
      * local rotation composed Ry, then Rx, then Rz, left-multiplied      (build.py:54-59)
      * angle index = position of the part in the dofs dict               (build.py:51-62)
      * while walking links: rot[child] = rot[child] @ rot[parent], but
        rot[child+'_i'] = rot[child+'_i'].T (transpose of the *local* transpose,
        toggling every time the child reappears)                           (build.py:78-79)
      * a child listed twice is overwritten                                (build.py:80)
      * output rows in pose_dict insertion order                           (build.py:82-86)
    """
    links = skel_dict["links"]
    positions = skel_dict["positions"]
    dofs = {k: list(v) for k, v in skel_dict["dofs"].items()}
    for joint in skel_dict["markers"]:
        dofs[joint] = [1, 1, 1]
    L = len(positions)
    parts = list(dofs.keys())

    def pose_to_3d(x):
        x = np.asarray(x, dtype=np.float64)
        phi, theta, psi = x[3:3 + L], x[3 + L:3 + 2 * L], x[3 + 2 * L:3 + 3 * L]
        rot = {}
        for i, part in enumerate(parts):
            Rm = np.eye(3)
            if dofs[part][1]:
                Rm = rot_y(theta[i]) @ Rm
            if dofs[part][0]:
                Rm = rot_x(phi[i]) @ Rm
            if dofs[part][2]:
                Rm = rot_z(psi[i]) @ Rm
            rot[part] = Rm
            rot[part + "_i"] = Rm.T
        pose = {}
        root = x[0:3]
        for link in links:
            if len(link) == 1:
                pose[link[0]] = root.copy()
                continue
            a, b = link
            if a not in pose:
                pose[a] = root.copy()
            tv = np.asarray(positions[b], dtype=np.float64) - np.asarray(positions[a], dtype=np.float64)
            rot[b] = rot[b] @ rot[a]
            rot[b + "_i"] = rot[b + "_i"].T
            pose[b] = pose[a] + rot[a + "_i"] @ tv
        return np.stack([pose[k] for k in pose], axis=0), list(pose.keys())

    return pose_to_3d
