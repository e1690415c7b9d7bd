#!/usr/bin/env python
"""Device-resident sequence replay (nav_frontend_sequence_dev, one k_frame_match launch per frame, programmatic
dependent launch): us per frame for one sequence and for eight side by side.  Used for A/B timing of kernel
variants: NAVSLAM_LIB=path/to/libnavslam_b200.so python profiles/prof_frame.py
usage: prof_frame.py [frames_per_repetition] [repetitions]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nav = importlib.import_module("nav-slam_b200")
synth = nav.synth
R, C_ = 64, 2048
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 30
NRES = 120   # 120 x 3.1 MB = 377 MB resident (> 126 MB L2)
frames = torch.from_numpy(synth.room_sequence(R, C_, NRES)).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for n_seq in (1, 8):
    ctx = nav.Context(R, C_, device=0, n_seq=n_seq)
    ctx.set_stream(stream.cuda_stream)
    if n_seq == 1:
        d = frames
        nf = NRES
    else:   # sequence s = the same room shifted in time: frame f of sequence s is frame (f + s) of the base
        nf = NRES - n_seq
        d = torch.stack([frames[s:s + nf] for s in range(n_seq)], dim=1).contiguous()   # [nf, n_seq, R, C, 3]
    pose = lambda f: np.array([50.0 * f, 0, 0, 0, 0, 0], dtype=np.float64)
    def poses(f0, n):
        pl = np.stack([[pose(f - 1 + s) for s in range(n_seq)] for f in range(f0, f0 + n)]).reshape(-1, 6)
        pf = np.stack([[pose(f + s) for s in range(n_seq)] for f in range(f0, f0 + n)]).reshape(-1, 6)
        return pf + np.array([-2.0, 0.5, 0, 0, 0, 0]), pl, pf
    fb = d.element_size() * d[0].numel()
    times = []
    n = min(F, nf - 1)
    pp, pl, pf = poses(1, n)
    evs = []
    for rep in range(REPS + 10):   # all repetitions queued back to back (the device never idles), one sync at the end
        ctx.slam_init_dev(d[0].data_ptr(), np.stack([pose(s) for s in range(n_seq)]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.frontend_sequence_dev(d.data_ptr() + fb, n, pp, pl, pf)
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    times = [a.elapsed_time(b) * 1e3 / n for a, b in evs[10:]]
    t = float(np.median(times))
    print(f"n_seq={n_seq}: {t:7.2f} us per step of {n_seq} frame(s) = {n_seq / t * 1e6:9.0f} frames/s "
          f"(min {min(times):.2f}, max {max(times):.2f}, {n} frames x {REPS} repetitions)", flush=True)
    ctx.close()
