"""Worker for tests/test_sharding_cpu.py: world_size-2 gloo run of the N>1 host logic."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sharding = importlib.import_module("nav-slam_b200.sharding")
    synth = importlib.import_module("nav-slam_b200.synth")
    from oracle_lib import Oracle
    oracle = Oracle()  # the checker stands in for the GPU kernel: this test is about the plumbing

    # replicated map: only rank 0 knows the points before the broadcast
    n, nq = 3000, 1001  # nq not divisible by world: uneven shards + padding
    pts = torch.zeros((n, 3), dtype=torch.float64)
    if rank == 0:
        pts = torch.from_numpy(synth.map_points(n, seed=5))
    sharding.broadcast_points(pts)
    q = torch.from_numpy(synth.map_queries(synth.map_points(n, seed=5), nq, seed=6))
    calls = []

    def nn_fn(qs):
        calls.append(int(qs.shape[0]))
        i, d = oracle.nn_brute(pts.numpy(), qs.numpy())
        return torch.from_numpy(i), torch.from_numpy(d)

    idx, dd = sharding.sharded_nn(nn_fn, q)
    lo, hi = sharding.shard_bounds(nq, world, rank)
    assert calls == [hi - lo], calls
    ref_i, ref_d = oracle.nn_brute(pts.numpy(), q.numpy())
    assert np.array_equal(idx.numpy(), ref_i) and np.array_equal(dd.numpy(), ref_d)

    # the in-place exchange of the hot loop: equal shards (two all_gathers, no packing) and the ragged case
    for nq2 in (1000, 1001):
        q2 = q[:nq2].contiguous()
        idx_full = torch.full((nq2,), -7, dtype=torch.int32)
        dist_full = torch.full((nq2,), -7.0, dtype=torch.float64)
        seen = []

        def nn_into(qs, iv, dv):
            seen.append(int(qs.shape[0]))
            i, d = oracle.nn_brute(pts.numpy(), qs.numpy())
            iv.copy_(torch.from_numpy(i))
            dv.copy_(torch.from_numpy(d))

        sharding.sharded_nn_into(nn_into, q2, idx_full, dist_full)
        a, b = sharding.shard_bounds(nq2, world, rank)
        assert seen == [b - a], seen
        assert np.array_equal(idx_full.numpy(), ref_i[:nq2]) and np.array_equal(dist_full.numpy(), ref_d[:nq2])

    # independent sequences: every rank processes its own, results meet on the host
    mine = sharding.assign_sequences(5, world)[rank]
    local = [(s, np.full(6, float(s))) for s in mine]
    merged = sharding.gather_sequence_results(local, world)
    assert sorted(merged) == [0, 1, 2, 3, 4] and all(merged[s][0] == s for s in merged)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
