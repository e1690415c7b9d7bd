#!/usr/bin/env python
"""Small driver for profiling / timing the batched curvature stencil (nav_extract_feature_batch_dev)
on a device-resident batch of 64x2048 images.  usage: prof_stencil.py [n_images] [reps]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
R, C = 64, 2048
base = nav.synth.room_sequence(R, C, 8)
d = torch.from_numpy(base).cuda().repeat((n + 7) // 8, 1, 1, 1)[:n].contiguous()
lab = torch.empty((n, R, C), dtype=torch.int32, device="cuda")
ctx = nav.Context(R, C, device=0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
for _ in range(3):
    ctx.extract_feature_batch_dev(d.data_ptr(), n, lab.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(reps):
    ctx.extract_feature_batch_dev(d.data_ptr(), n, lab.data_ptr())
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
b = n * R * C * 28
print(f"labels_batch n={n}: {ms*1e3:.1f} us/launch, {b/ms/1e6:.1f} GB/s algorithmic "
      f"({b/ms/1e6/6534.5:.3f} of measured HBM peak), exact fallbacks={ctx.exact_fallback_count()}")
