"""``lib.misc`` of the reference's newer API (the package is missing from the snapshot; names and semantics from its call
sites in /root/reference/src/all_optimizations.py:497,583,584,618,883,927)."""
import numpy as np

from .build import redescending_loss  # noqa: F401  (all_optimizations.py:497; literal formula of build.py:382-395)
from .ekf import get_3d_marker_coords, get_pose_params  # noqa: F401
from .fte import MARKERS


def get_markers():
    """The 20 cheetah markers in the row order of ``positions`` (all_optimizations.py:170-178)."""
    return list(MARKERS)


def rot_x(x):
    """all_optimizations.py:66-73 (NumPy)"""
    c, s = np.cos(x), np.sin(x)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])


def rot_y(y):
    """all_optimizations.py:75-82"""
    c, s = np.cos(y), np.sin(y)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])


def rot_z(z):
    """all_optimizations.py:84-91"""
    c, s = np.cos(z), np.sin(z)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
