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


def _run_world2(worker, port):
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", worker)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=280)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


@pytest.mark.timeout(300)
def test_interface_gather_world2_gloo():
    _run_world2("_rank_worker.py", 29500 + (os.getpid() % 2000))


@pytest.mark.timeout(300)
def test_sba_sharded_camera_system_world2_gloo():
    """Sharded SBA: per-rank Schur complement + ONE all_reduce of [S | rhs] == the global step."""
    _run_world2("_rank_worker_sba.py", 31600 + (os.getpid() % 2000))


def test_shard_points_plan():
    from acinoset_b200 import sba

    pidx = np.repeat(np.arange(10), 3)
    for w in (1, 2, 3, 8):
        plan = sba.shard_points(pidx, 10, w)
        assert len(plan) == w and sum(n for _, n, _ in plan) == 10
        assert np.array_equal(np.sort(np.concatenate([i for _, _, i in plan])), np.arange(30))
        for p0, n, ids in plan:
            assert np.all((pidx[ids] >= p0) & (pidx[ids] < p0 + n))
