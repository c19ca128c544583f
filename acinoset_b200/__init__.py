"""acinoset_b200 - B200 (sm_100a) implementation of AcinoSet's reprojection /
trajectory-optimisation hot path behind the reference's own Python entry points.

The CUDA extension (libacino_b200.so) is required; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the extension is missing)
from ._lib import AcinoError, Handle  # noqa: F401

__version__ = "0.1.0"
