#!/usr/bin/env python
"""Where the time of one frame goes inside k_frame_seq: needs a library built with -DNAV_SEQ_TIMING
(developer instrumentation: thread 0 of every CTA stamps %globaltimer at seven points of every frame).
  NAVSLAM_LIB=<variant .so> python profiles/prof_seq_phases.py
Prints, averaged over CTAs and frames, the duration of each phase and the time CTAs spend waiting for their
row's cluster.  Build the variant with
  nvcc <flags of nav-slam_b200/build.py> -DNAV_SEQ_TIMING -c csrc/rowmap.cu  and link it with the other objects."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")
L = nav.load_library()
R, C_, F = 64, 2048, 60
frames = torch.from_numpy(nav.synth.room_sequence(R, C_, F + 1)).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = nav.Context(R, C_, device=0)
ctx.set_stream(stream.cuda_stream)
n_cta = R * (C_ // 256)
stamps = torch.zeros((F, n_cta, 8), dtype=torch.int64, device="cuda")
L.nav_debug_set_seq_stamps.argtypes = [C.c_void_p]
pose = lambda f: np.array([50.0 * f, 0, 0, 0, 0, 0], dtype=np.float64)
pl = np.stack([pose(f - 1) for f in range(1, F + 1)])
pf = np.stack([pose(f) for f in range(1, F + 1)])
pp = pf + np.array([-2.0, 0.5, 0, 0, 0, 0])
for rep in range(3):
    ctx.slam_init_dev(frames[0].data_ptr(), pose(0))
    if rep == 2:
        assert L.nav_debug_set_seq_stamps(stamps.data_ptr()) == 0
    ctx.frontend_sequence_dev(frames[1].data_ptr(), F, pp, pl, pf)
    torch.cuda.synchronize()
t = stamps.cpu().numpy().astype(np.float64)[5:]          # skip the pipeline fill
names = ["wait for the tile", "labels (fp32 filter)", "wait for the row's cluster (previous frame)",
         "map prefetch + next-tile issue + label store + map_tile", "wait neighbourhood + compaction", "search"]
d = np.diff(t[:, :, :7], axis=2)                           # [frames, ctas, 6] ns
print(f"k_frame_seq phases, ns, mean over {d.shape[0]} frames x {d.shape[1]} CTAs (thread 0 of each CTA; min / mean / max over CTAs of the per-CTA mean)")
for k, nme in enumerate(names):
    per_cta = d[:, :, k].mean(axis=0)
    print(f"  {nme:58s} {per_cta.min():8.0f} {per_cta.mean():8.0f} {per_cta.max():8.0f}")
frame_t = (t[1:, :, 0] - t[:-1, :, 0]).mean()
print(f"  frame to frame (stamp 0 to stamp 0) {frame_t:8.0f} ns")
ctx.close()
# which tiles / rows are the slow ones (search phase, ns, mean over frames)
n_tiles = C_ // 256
srch = d[:, :, 5].mean(axis=0).reshape(R, n_tiles)
print("search ns per tile (mean over rows):   " + " ".join(f"{v:6.0f}" for v in srch.mean(axis=0)))
print("search ns per row (mean over tiles), rows 0,8,..: " + " ".join(f"{v:6.0f}" for v in srch.mean(axis=1)[::8]))
print("slowest tile of a row (mean over rows): %.0f   mean tile: %.0f" % (srch.max(axis=1).mean(), srch.mean()))
