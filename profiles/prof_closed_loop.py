#!/usr/bin/env python
"""Closed loop with pose feedback as ONE C call (nav_slam_run, 64x2048): wall time per frame for xyz and
depth input, and with NAV_RUN_TRACE=1 the host-side split (prefetch call / launches / wait for the
statistics / fit / mapping call).  usage: prof_closed_loop.py [n_frames]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")
synth = nav.synth

R, C_ = 64, 2048
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
NRES = 40
clouds = synth.room_sequence(R, C_, NRES + 1)
frames = torch.from_numpy(clouds).pin_memory()
depth = torch.from_numpy(np.stack([synth.l5_depth_frame(f, R, C_) for f in range(8)])).pin_memory()
ctx = nav.Context(R, C_)
step = np.array([48.0, 0.5, 0.0, 0.0, 0.0, 0.0])
for use_depth in (False, True):
    src = depth if use_depth else frames
    n_src = src.shape[0]
    ptrs = [src[1 + (f % (n_src - 1))].data_ptr() for f in range(N)]
    deltas = np.tile(np.array([-19.0, 0.3, 0, 0, 0, 0]) if use_depth else step, (N, 1))
    for rep in range(3):
        ctx.slam_init(np.zeros(6), clouds[0], want_global=False)
        ctx.synchronize()
        t0 = time.perf_counter()
        poses, errs, ncs = ctx.slam_run(ptrs, deltas, np.zeros(6), depth_input=use_depth)
        ctx.synchronize()
        t = (time.perf_counter() - t0) / N
        print(f"nav_slam_run {'depth' if use_depth else 'xyz  '} input: {t*1e6:7.1f} us/frame = {1/t:8.0f} frames/s "
              f"({N} frames, {int(ncs[-1])} correspondences, rms {errs[-1]:.2f} mm)", flush=True)
ctx.close()
