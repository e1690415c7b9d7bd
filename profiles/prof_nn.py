#!/usr/bin/env python
"""Large-map NN: flat kd-tree (k_kd_nn) against the exact fp64 brute-force kernel (k_bf_nn) across
map sizes, 131 072 jittered queries, device resident.  Prints one line per size; used for the
tree-vs-brute-force decision recorded in DESIGN.md and for the ncu capture of k_kd_nn.
usage: prof_nn.py [sizes...]   (default 1024 4096 16384 65536 1000000)"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")

sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 16384, 65536, 1_000_000]
NQ = 131072
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
s = stream.cuda_stream


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in sizes:
    pts = nav.synth.map_points(n, seed=n)
    q = nav.synth.map_queries(pts, NQ, seed=n + 1)
    if os.environ.get("PRESORT"):  # host-side Morton order, to time the traversal on coherent queries
        lo, hi = pts.min(0), pts.max(0)
        g = np.clip(((q - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023)
        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v
        key = spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)
        q = np.ascontiguousarray(q[np.argsort(key, kind="stable")])
    d_pts, d_q = torch.from_numpy(pts).cuda(), torch.from_numpy(q).cuda()
    i1 = torch.empty(NQ, dtype=torch.int32, device="cuda")
    d1 = torch.empty(NQ, dtype=torch.float64, device="cuda")
    i2, d2 = torch.empty_like(i1), torch.empty_like(d1)
    tree = nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=0, stream=s)
    build_ms = timed(lambda: nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=0, stream=s).close(), reps=3)
    kd_ms = timed(lambda: tree.nn_batch_dev(d_q.data_ptr(), NQ, i1.data_ptr(), d1.data_ptr(), s))
    bf_ms = None
    if n <= 65536:
        bf_ms = timed(lambda: nav.bruteforce_nn_dev(0, d_pts.data_ptr(), n, d_q.data_ptr(), NQ, i2.data_ptr(),
                                                    d2.data_ptr(), stream=s), reps=2)
        assert torch.equal(i1, i2) and torch.equal(d1, d2), "kd-tree and brute force disagree"
    print(f"n={n:8d}  kd build {build_ms:7.3f} ms  kd query {kd_ms*1e3:8.1f} us ({NQ/kd_ms/1e3:7.1f} Mq/s)  "
          + (f"fp64 brute force {bf_ms*1e3:9.1f} us ({n*NQ/bf_ms/1e6:6.1f} Gpair/s)  same answers" if bf_ms else ""))
    tree.close()
