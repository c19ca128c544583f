import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def dummy_cams():
    import synth

    return synth.load_dummy_scene()


@pytest.fixture(scope="session")
def fte_problem_small(dummy_cams):
    """Seeded 64-frame synthetic FTE problem generated with the fp64 oracle."""
    import synth
    from oracle import fisheye, skeleton

    return synth.make_fte_problem(64, skeleton.cheetah_fk_active, fisheye.project, seed=11, cams=dummy_cams)
