"""Oracle helper (test infrastructure): ctypes loader of oracle/liboracle_c.so, the plain-C
fp64 restatement in oracle/c/fte_oracle.c (OpenMP over frames).  Used by tests (validated
against the NumPy oracle) and as the CPU baseline / --impl reference arm of bench.py."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        _lib.fte_oracle_eval.argtypes = [ci, ci] + [vp] * 11 + [ci]
        _lib.fte_oracle_eval.restype = None
        _lib.fte_oracle_max_threads.restype = ci
        _lib.fte_oracle_fk.argtypes = [ci, vp, vp]
    return _lib


def max_threads():
    return int(load().fte_oracle_max_threads())


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def fte_eval(xa, meas, w, K, D, R, t, abc=(3.0, 10.0, 20.0), want_H=True, n_threads=0):
    lib = load()
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    xa, meas, w = c(xa), c(meas), c(w)
    N, C = xa.shape[0], meas.shape[1]
    K, D, R, t = c(K).reshape(C, 9), c(D).reshape(C, 4), c(R).reshape(C, 9), c(t).reshape(C, 3)
    abc = c(abc)
    cost = np.empty(N)
    g = np.empty((N, 25))
    H = np.empty((N, 25, 25)) if want_H else None
    lib.fte_oracle_eval(N, C, _p(xa), _p(meas), _p(w), _p(K), _p(D), _p(R), _p(t), _p(abc), _p(cost), _p(g), _p(H),
                        int(n_threads))
    return cost, g, H


def cheetah_fk(xa):
    lib = load()
    xa = np.ascontiguousarray(xa, dtype=np.float64)
    pos = np.empty((xa.shape[0], 20, 3))
    lib.fte_oracle_fk(xa.shape[0], _p(xa), _p(pos))
    return pos
