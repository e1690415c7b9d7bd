#!/usr/bin/env python
"""PCIe copy rates for the frame sizes of the e2e path (pinned memory, torch streams)."""
import time
import torch

n_in, n_out = 64 * 2048 * 24, 64 * 2048 * 16
h_in = torch.empty(n_in * 8, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out * 8, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in * 8, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out * 8, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=400):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        k = i % 8
        if h2d:
            with torch.cuda.stream(s1):
                d_in[k * n_in:(k + 1) * n_in].copy_(h_in[k * n_in:(k + 1) * n_in], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[k * n_out:(k + 1) * n_out].copy_(d_out[k * n_out:(k + 1) * n_out], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, a, b in (("H2D 3.1 MB alone", True, False), ("D2H 2.1 MB alone", False, True), ("both directions", True, True)):
    run(a, b, 50)
    dt = run(a, b)
    print(f"{name:18s} {dt*1e6:7.1f} us per frame   H2D {n_in/dt/1e9 if a else 0:5.1f} GB/s   D2H {n_out/dt/1e9 if b else 0:5.1f} GB/s")
