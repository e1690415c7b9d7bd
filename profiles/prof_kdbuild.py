#!/usr/bin/env python
"""kd build time per call (CUDA events on the build stream), device-resident points.
usage: prof_kdbuild.py [n ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
s = stream.cuda_stream
args = sys.argv[1:]
for a in args or ["65536", "1000000", "10000000", "room"]:
    if a == "room":   # config 4 as a SLAM run produces it: 8 mapped 64x2048 room frames (points on planes)
        pts, _ = nav.synth.accumulated_map(8)
        n = pts.shape[0]
        d_pts = torch.from_numpy(pts).cuda()
    else:
        n = int(a)
        d_pts = torch.from_numpy(nav.synth.map_points(n, seed=n)).cuda()
    for split in ("widest", "cyclic"):
        times = []
        for rep in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t = nav.KdTree(dev_ptr=d_pts.data_ptr(), n=n, device=0, stream=s, split=split)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            launches = t.launch_count()
            t.close()
        label = "room map" if a == "room" else f"n={n}"
        print(f"{label:>12s} split={split:6s} build ms per call: " + " ".join(f"{x:6.2f}" for x in times) + f"   launches {launches}")
