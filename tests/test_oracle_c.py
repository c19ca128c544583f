"""The C restatement (oracle/c/fte_oracle.c, the CPU baseline) against the NumPy oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cport():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    from oracle import c_port

    return c_port


def test_c_port_fk(cport):
    from oracle import skeleton

    rng = np.random.default_rng(0)
    xa = rng.normal(0, 1, (50, 25))
    assert np.abs(cport.cheetah_fk(xa) - skeleton.cheetah_fk_active(xa)).max() < 1e-13


def test_c_port_fte_eval(cport, fte_problem_small):
    from oracle import fte

    p = fte_problem_small
    K, D, R, t, _ = p["cams"]
    c_ref, g_ref, H_ref = fte.fte_eval(p["x0"], p["meas"], p["w"], K, D, R, t)
    for nt in (1, 0):
        c, g, H = cport.fte_eval(p["x0"], p["meas"], p["w"], K, D, R, t, n_threads=nt)
        assert np.abs(c - c_ref).max() < 1e-9 * np.abs(c_ref).max()
        assert np.abs(g - g_ref).max() < 1e-9 * np.abs(g_ref).max()
        assert np.abs(H - H_ref).max() < 1e-9 * np.abs(H_ref).max()
