"""Oracle helper (test infrastructure): import the UNMODIFIED reference modules from
/root/reference in the build container, to validate the restatement and to generate
golden vectors (tests/golden/make_golden.py).  /root/reference does not exist on the
GPU box; nothing that runs there may call this.

Three shims are installed before import:
  * a stub ``nptyping`` whose ``Array`` supports ``Array[...]``  (calib.py:3, utils.py:4)
  * ``np.float = float; np.int = int``                         (calib.py:261-262,409-410)
  * a stub ``matplotlib.pyplot``                                (calib.py:10)
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("ACINO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "calib"))


def _install_shims():
    import numpy as np

    if "nptyping" not in sys.modules:
        m = types.ModuleType("nptyping")

        class _Array:
            def __class_getitem__(cls, item):
                return cls

        m.Array = _Array
        sys.modules["nptyping"] = m
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load_reference_calib():
    """Returns (calib, utils) = the reference's src/calib/{calib,utils}.py, unmodified."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_shims()
    # load as a package named 'calib' WITHOUT running its __init__ siblings that need PyQt
    pkg_dir = os.path.join(REF_ROOT, "src", "calib")
    if "calib" not in sys.modules or getattr(sys.modules["calib"], "__path__", [None])[0] != pkg_dir:
        pkg = types.ModuleType("calib")
        pkg.__path__ = [pkg_dir]
        sys.modules["calib"] = pkg
    utils = importlib.import_module("calib.utils")
    calib = importlib.import_module("calib.calib")
    return calib, utils


def reference_source_lines(relpath, first, last):
    """Text of lines [first, last] (1-based, inclusive) of a reference source file."""
    with open(os.path.join(REF_ROOT, relpath)) as f:
        lines = f.readlines()
    return "".join(lines[first - 1:last])
