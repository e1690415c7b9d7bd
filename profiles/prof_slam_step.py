#!/usr/bin/env python
"""The whole SLAM step with pose feedback through the C ABI (64x2048): localisation from pinned host
frames -- labels, NN, per-row dedupe, sufficient statistics on the device, the 200-iteration Adam fit
in O(1) per iteration on the host -- then mapping with the fitted pose, and the same with the mapped
cloud downloaded.  Next to it the reference's slam_localization + slam_mapping on one core."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
nav = importlib.import_module("nav-slam_b200")
synth = nav.synth

R, C_ = 64, 2048
N = 40
frames = torch.from_numpy(synth.room_sequence(R, C_, N + 1)).pin_memory()
ctx = nav.Context(R, C_)
pos = np.zeros(6)
ctx.slam_init(pos, frames[0].numpy(), want_global=False)
for want_global in (False, True):
    ctx.slam_init(np.zeros(6), frames[0].numpy(), want_global=False)
    last = np.zeros(6)
    times = []
    for f in range(1, N + 1):
        pred = last + np.array([48.0, 0.5, 0.0, 0.0, 0.0, 0.0])          # motion model: a little off the truth
        t0 = time.perf_counter()
        p, err, ncorr = ctx.slam_localization_fast(frames[f].numpy(), pred, last)
        ctx.slam_mapping(p, None, want_global=want_global)
        times.append(time.perf_counter() - t0)
        last = p
    t = np.median(times[5:])
    print(f"GPU step (localization_fast + mapping{', global cloud downloaded' if want_global else ''}): "
          f"{t*1e6:8.1f} us/frame = {1/t:7.1f} frames/s   pose x after {N} frames {last[0]:.3f} mm (truth {50.0*N:.1f}), "
          f"{ncorr} correspondences, rms {err:.3f} mm")
try:
    from oracle_lib import RefLib, Pos, ref_available, quiet_stdout
    if ref_available(f"{R}x{C_}"):
        ref = RefLib(R, C_)
        attr = np.zeros(ref.sizeof_slam_attr, dtype=np.uint8)
        pc = [ref.pack_cloud(frames[f].numpy(), ts=f) for f in range(4)]
        ref.lib.init_slam(attr.ctypes.data, Pos.of(np.zeros(6)), pc[0].ctypes.data)
        last = np.zeros(6)
        ts = []
        with quiet_stdout():
            for f in range(1, 4):
                pred = last + np.array([48.0, 0.5, 0.0, 0.0, 0.0, 0.0])
                t0 = time.perf_counter()
                p = ref.lib.slam_localization(attr.ctypes.data, pc[f].ctypes.data, Pos.of(pred), Pos.of(last)).arr()
                ref.lib.slam_mapping(attr.ctypes.data, Pos.of(p), pc[f].ctypes.data)
                ts.append(time.perf_counter() - t0)
                last = p
        print(f"reference slam_localization + slam_mapping, one core: {np.median(ts)*1e3:8.1f} ms/frame = {1/np.median(ts):.2f} frames/s")
except Exception as e:  # noqa: BLE001
    print("reference step not timed:", e)
ctx.close()
