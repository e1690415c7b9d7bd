#!/usr/bin/env python
"""Small all-kernel pass for compute-sanitizer (memcheck / racecheck): 16x512 frames through the
blocking and pipelined frame paths, dedupe + fit, batch labels (TMA kernel), kd build + query."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")
R, C = 16, 512
ctx = nav.Context(R, C, device=0)
fr = [nav.synth.room_frame(R, C, f) for f in range(4)]
z = np.zeros(6)
ctx.slam_init(z, fr[0])
p = z + np.array([50.0, 0, 0, 0, 0, 0])
ctx.frontend_frame(fr[1], p, z, p)
ctx.slam_match(fr[2], 2 * p, p)
ctx.slam_localization_fast(fr[2], 2 * p, p)
ctx.slam_mapping(2 * p, None)
ctx.row_map_export(3)
ctx.curvature(fr[3])
ctx.convert_to_pointcloud(np.full((R, C), 1000, dtype=np.int32))
ctx.flatten_points(fr[0][0], np.ones(C, dtype=np.int32))
h = torch.from_numpy(fr[3]).pin_memory()
o = [torch.empty((R, C), dtype=torch.int32).pin_memory(), torch.empty((R, C), dtype=torch.int32).pin_memory(),
     torch.empty((R, C), dtype=torch.float64).pin_memory(), torch.empty((R, C, 3), dtype=torch.float64).pin_memory()]
ctx.frontend_frame_async(h.data_ptr(), 3 * p, 2 * p, 3 * p, *[t.data_ptr() for t in o])
ctx.frontend_wait()
batch = torch.from_numpy(np.stack([nav.synth.room_frame(R, C, f) for f in range(80)])).cuda()
lab = torch.empty((80, R, C), dtype=torch.int32, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.extract_feature_batch_dev(batch.data_ptr(), 80, lab.data_ptr())
torch.cuda.synchronize()
ctx.set_stream(None)
pts = nav.synth.map_points(20000, seed=3)
t = nav.KdTree(pts)
t.nn_batch(nav.synth.map_queries(pts, 9000, seed=4))
t.close()
ctx.close()
print("sanitize_smoke done")
