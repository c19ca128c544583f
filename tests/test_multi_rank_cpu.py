"""N > 1 host logic on CPU: world_size-2 gloo run of the interface gather / chain assembly /
scatter used by the multi-GPU FTE solve, and the frame sharding plan."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_plan():
    from acinoset_b200 import lm

    for n, w in [(100000, 8), (10000, 8), (1000, 2), (90, 4), (7, 2)]:
        sh = lm.shard_frames(n, w)
        assert len(sh) == w and sh[0][0] == 0 and sum(c for _, c in sh) == n
        for r, (f0, c) in enumerate(sh):
            assert f0 == sum(cc for _, cc in sh[:r])
            if r < w - 1 and c > 0:
                assert c % 3 == 0
    assert lm.shard_frames(100000, 8)[0] == (0, 12501)


def test_default_bounds_match_reference_table():
    from acinoset_b200 import lm
    from oracle import skeleton

    lo, hi = lm.default_bounds()
    lo2, hi2 = skeleton.active_bounds()
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    assert np.array_equal(lm.Q_SIGMA_ACTIVE, skeleton.Q_SIGMA[skeleton.ACTIVE_IDX])


@pytest.mark.timeout(300)
def test_interface_gather_world2_gloo():
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_rank_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=280)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
