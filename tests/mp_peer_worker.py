"""Worker for tests/test_peer_gpu.py (needs >= 2 GPUs): the peer-memory exchange of config 5b (nav_peer_*,
sharding.PeerGather) against the oracle's brute force and against the NCCL all_gather path, ragged shards,
several calls in a row (double buffering), world_size ranks under torchrun."""
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
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nav = importlib.import_module("nav-slam_b200")
    sharding = importlib.import_module("nav-slam_b200.sharding")
    from oracle_lib import Oracle
    oracle = Oracle()
    L = nav.load_library()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    s = stream.cuda_stream
    n = 20000
    pts = nav.synth.map_points(n, seed=5)
    d_pts = torch.from_numpy(pts).cuda()
    tree = nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=local, stream=s)
    for nq in (4001, 4096):                       # ragged and even shards
        pg = sharding.PeerGather(L, local, nq)
        for call in range(5):                     # more calls than buffer halves
            q = nav.synth.map_queries(pts, nq, seed=100 + call)
            d_q = torch.from_numpy(q).cuda()
            idx, dd = pg.nn(tree, d_q, s)
            torch.cuda.synchronize()
            oi, od = oracle.nn_brute(pts, q)
            assert np.array_equal(idx.cpu().numpy(), oi), (rank, nq, call)
            assert np.array_equal(dd.cpu().numpy(), od), (rank, nq, call)
        pg.check()
        # the NCCL path gives the same arrays
        bi = torch.empty(nq, dtype=torch.int32, device="cuda")
        bd = torch.empty(nq, dtype=torch.float64, device="cuda")

        def nn_into(qs, iv, dv):
            tree.nn_batch_dev(qs.data_ptr(), int(qs.shape[0]), iv.data_ptr(), dv.data_ptr(), s)
        sharding.sharded_nn_into(nn_into, d_q, bi, bd)
        torch.cuda.synchronize()
        assert torch.equal(bi, idx) and torch.equal(bd, dd)
        # the MAP sharded: every rank builds a tree over its part of the points; same answers, also with clustered
        # points and exact duplicates across the shard boundary (the merge breaks ties on the global index)
        for variant in ("uniform", "dup"):
            mp = pts if variant == "uniform" else np.repeat(pts[: n // 4], 4, axis=0)
            lo, hi = sharding.shard_bounds(mp.shape[0], world, rank)
            d_part = torch.from_numpy(np.ascontiguousarray(mp[lo:hi])).cuda()
            part = nav.KdTree(dev_ptr=d_part.data_ptr(), n=hi - lo, device=local, stream=s)
            for call in range(3):
                q = nav.synth.map_queries(mp, nq, seed=200 + call)
                if variant == "dup":
                    q[: nq // 2] = mp[np.random.default_rng(call).integers(0, mp.shape[0], nq // 2)]   # exact hits: ties
                d_q2 = torch.from_numpy(q).cuda()
                idx, dd = pg.nn_sharded_map(part, d_q2, lo, s)
                torch.cuda.synchronize()
                oi, od = oracle.nn_brute(mp, q)
                assert np.array_equal(idx.cpu().numpy(), oi), (rank, nq, variant, call)
                assert np.array_equal(dd.cpu().numpy(), od), (rank, nq, variant, call)
            pg.check()
            dist.barrier()
            part.close()
        dist.barrier()
        pg.close()
    tree.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
