#!/usr/bin/env python
"""Config 4 as a SLAM run produces it: an accumulated map (8 mapped 64x2048 frames = 1 048 576 points)
queried by the next frame's 131 072 points in image order, against the same map queried in random
order.  Shows what query coherence is worth to each traversal kernel (NAV_KD_KERNEL)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
s = stream.cuda_stream


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


pts, q_img = nav.synth.accumulated_map(int(os.environ.get("FRAMES", "8")))
n, nq = pts.shape[0], q_img.shape[0]
perm = np.random.default_rng(5).permutation(nq)
q_rnd = np.ascontiguousarray(q_img[perm])
d_pts = torch.from_numpy(pts).cuda()
tree = nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=0, stream=s)
res = {}
for name, q in (("image order", q_img), ("random order", q_rnd)):
    d_q = torch.from_numpy(q).cuda()
    i1 = torch.empty(nq, dtype=torch.int32, device="cuda")
    d1 = torch.empty(nq, dtype=torch.float64, device="cuda")
    ms = timed(lambda: tree.nn_batch_dev(d_q.data_ptr(), nq, i1.data_ptr(), d1.data_ptr(), s))
    res[name] = (i1.cpu().numpy(), d1.cpu().numpy())
    print(f"map {n} pts (accumulated room), {nq} queries in {name:12s}: {ms*1e3:8.1f} us  {nq/ms/1e3:8.1f} Mq/s"
          f"   kernel={os.environ.get('NAV_KD_KERNEL', 'default')}")
a, b = res["image order"], res["random order"]
assert np.array_equal(a[0][perm], b[0]) and np.array_equal(a[1][perm], b[1])
print("same answers in both orders; mean NN distance %.1f mm" % a[1].mean())
tree.close()
