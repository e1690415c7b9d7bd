"""N>1 host logic on CPU: shard arithmetic, and a world_size-2 gloo run of the query sharding /
sequence partitioning plumbing (nav-slam_b200/sharding.py)."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sharding = importlib.import_module("nav-slam_b200.sharding")


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 128, 131072, 1000001):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.max_shard(n, world)
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_sequence_assignment():
    assert sharding.assign_sequences(8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert sharding.assign_sequences(8, 8) == [[i] for i in range(8)]
    assert sharding.assign_sequences(3, 4) == [[0], [1], [2], []]


def test_world_size_2_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mp_sharding_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout


def test_library_shard_arithmetic_agrees_with_python():
    """The peer-memory kernels route answers with nav_shard_range / nav_shard_owner (csrc/nav_kdtree.cuh); the Python
    side cuts the queries with sharding.shard_bounds.  Host functions: no GPU needed."""
    import ctypes as C
    nav = importlib.import_module("nav-slam_b200")
    L = nav.load_library()
    for n in (0, 1, 7, 8, 9, 1001, 131072, 131073):
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                L.nav_shard_range(n, world, r, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == sharding.shard_bounds(n, world, r)
                for i in {lo.value, (lo.value + hi.value) // 2, hi.value - 1}:
                    if lo.value <= i < hi.value:
                        assert L.nav_shard_owner(n, world, i) == r
