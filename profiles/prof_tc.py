#!/usr/bin/env python
"""Tensor-core brute force (tcgen05 candidate tiles + exact re-rank) against the exact fp64 brute-force
kernel and the kd-tree: correctness (bit-identical idx/dist) and time, 131 072 queries by default.
usage: prof_tc.py [nq] [sizes...]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
sizes = [int(a) for a in sys.argv[2:]] or [256, 1024, 4096, 16384, 65536]
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
    for variant in ("uniform", "clustered"):
        pts = nav.synth.map_points(n, variant=variant, seed=n)
        q = nav.synth.map_queries(pts, NQ, seed=n + 1)
        d_pts, d_q = torch.from_numpy(pts).cuda(), torch.from_numpy(q).cuda()
        out = [(torch.empty(NQ, dtype=torch.int32, device="cuda"), torch.empty(NQ, dtype=torch.float64, device="cuda"))
               for _ in range(3)]
        tree = nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=0, stream=s)
        kd = timed(lambda: tree.nn_batch_dev(d_q.data_ptr(), NQ, out[0][0].data_ptr(), out[0][1].data_ptr(), s))
        tc = timed(lambda: nav.bruteforce_nn_dev(0, d_pts.data_ptr(), n, d_q.data_ptr(), NQ, out[1][0].data_ptr(),
                                                 out[1][1].data_ptr(), use_tensor_cores=True, stream=s))
        bf = None
        if n * NQ <= 2 ** 34:
            bf = timed(lambda: nav.bruteforce_nn_dev(0, d_pts.data_ptr(), n, d_q.data_ptr(), NQ, out[2][0].data_ptr(),
                                                     out[2][1].data_ptr(), stream=s), reps=2)
            ok_bf = torch.equal(out[1][0], out[2][0]) and torch.equal(out[1][1], out[2][1])
        ok_kd = torch.equal(out[1][0], out[0][0]) and torch.equal(out[1][1], out[0][1])
        print(f"n={n:7d} {variant:9s} tensor-core {tc*1e3:8.1f} us | kd-tree {kd*1e3:7.1f} us | fp64 brute "
              + (f"{bf*1e3:8.1f} us" if bf else "   --   ") + f" | tc==kd {ok_kd}" + (f" tc==bf {ok_bf}" if bf else ""),
              flush=True)
        tree.close()
