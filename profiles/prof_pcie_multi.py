#!/usr/bin/env python
"""Concurrent host<->device copy probe over all ranks of one node (torchrun --nproc-per-node N): every rank
copies pinned frames of the e2e sizes to and from its own GPU at the same time; prints per-rank and aggregate
rates.  Names the host-side ceiling the multi-GPU e2e numbers run into.
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/prof_pcie_multi.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

SIZES = {"xyz in 3.1 MB / masks+nn out 1.6 MB": (64 * 2048 * 24, 64 * 2048 * 12 + 64 * 128 * 4),
         "depth in 0.5 MB / masks+nn out 1.6 MB": (64 * 2048 * 4, 64 * 2048 * 12 + 64 * 128 * 4)}
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(n_in, n_out, h2d, d2h, reps):
    h_in = torch.empty(n_in * 8, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out * 8, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in * 8, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out * 8, dtype=torch.uint8, device="cuda")

    def loop(n):
        for i in range(n):
            k = i % 8
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[k * n_in:(k + 1) * n_in].copy_(h_in[k * n_in:(k + 1) * n_in], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[k * n_out:(k + 1) * n_out].copy_(d_out[k * n_out:(k + 1) * n_out], non_blocking=True)
    loop(50)
    barrier()
    t0 = time.perf_counter()
    loop(reps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt


for name, (n_in, n_out) in SIZES.items():
    for label, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both", True, True)):
        dt = run(n_in, n_out, a, b, 600)
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            all_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            dts = [float(x) for x in all_t]
        else:
            dts = [dt]
        if rank == 0:
            worst = max(dts)
            agg_in = sum(n_in / d for d in dts) / 1e9 if a else 0.0
            agg_out = sum(n_out / d for d in dts) / 1e9 if b else 0.0
            print(f"N={world} {name:40s} {label:9s}: slowest rank {worst*1e6:7.1f} us/frame "
                  f"({1/worst:8.0f} frames/s per rank, {world/worst:9.0f} aggregate)   aggregate H2D {agg_in:6.1f} GB/s  "
                  f"D2H {agg_out:6.1f} GB/s   per-rank us: " + " ".join(f"{d*1e6:.0f}" for d in dts), flush=True)
if world > 1:
    dist.destroy_process_group()
