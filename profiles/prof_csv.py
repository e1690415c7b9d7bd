"""GPU CSV row writer: nav_csv_format_frame_gpu on the resident global cloud, text into
pinned memory, against the host writer and the printf restatement.  Usage: python profiles/prof_csv.py"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle_lib import Oracle  # noqa: E402

nav = importlib.import_module("nav-slam_b200")
synth = importlib.import_module("nav-slam_b200.synth")


def main():
    orc = Oracle()
    L = nav.load_library()
    for (rows, cols) in ((16, 1800), (64, 2048)):
        cloud = synth.room_frame(rows, cols, 0)
        pos = np.array([12.5, -3.25, 100.0, 0.1, 0.2, 0.3])
        ctx = nav.Context(rows, cols)
        g = ctx.slam_init(pos, cloud)
        cap = rows * cols * 700
        pin = L.nav_host_alloc(cap)
        n = ctx.csv_rows(7, pos, ekf_pos=pos, out_ptr=pin, out_cap=cap)
        text = C.string_at(pin, n)
        t0 = time.perf_counter()
        want = orc.csv_format_frame(7, g, pos, ekf_pos=pos)
        t_printf = time.perf_counter() - t0
        assert text == want
        for _ in range(5):
            ctx.csv_rows(7, pos, ekf_pos=pos, out_ptr=pin, out_cap=cap)
        reps = 50
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.csv_rows(7, pos, ekf_pos=pos, out_ptr=pin, out_cap=cap)
        t_gpu = (time.perf_counter() - t0) / reps
        # device-resident variant: kernels only (CUDA events around the call are on another stream; use wall
        # time of the synchronising call, which has no text download)
        dtext = torch.empty(cap, dtype=torch.uint8, device="cuda")
        nb = C.c_size_t(0)
        lp = nav.binding._pos_array(pos)
        for _ in range(5):
            L.nav_csv_format_frame_dev(ctx.h, 7, None, None, None, lp, lp, dtext.data_ptr(), cap, C.byref(nb))
        t0 = time.perf_counter()
        for _ in range(reps):
            L.nav_csv_format_frame_dev(ctx.h, 7, None, None, None, lp, lp, dtext.data_ptr(), cap, C.byref(nb))
        t_dev = (time.perf_counter() - t0) / reps
        assert nb.value == n and bytes(dtext[:n].cpu().numpy()) == want
        buf = C.create_string_buffer(cap)
        gp = np.ascontiguousarray(g)
        t_host = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            m = L.nav_csv_format_frame(buf, cap, 7, rows, cols, gp.ctypes.data, None, None, lp, lp)
            t_host = min(t_host, time.perf_counter() - t0)
        assert m == n
        mb = n / 1e6
        print(f"{rows}x{cols}: {mb:.2f} MB of CSV per frame ({n / (rows * cols):.0f} B/line)")
        print(f"  nav_csv_format_frame_gpu (resident cloud -> pinned text)  {t_gpu*1e6:9.1f} us  {mb/t_gpu/1e3:7.2f} GB/s")
        print(f"  nav_csv_format_frame_dev (text stays in HBM, incl. sync)  {t_dev*1e6:9.1f} us  {mb/t_dev/1e3:7.2f} GB/s")
        print(f"  nav_csv_format_frame     (host, one core)                 {t_host*1e6:9.1f} us  {mb/t_host/1e3:7.2f} GB/s")
        print(f"  snprintf restatement of main.c:324 (one core)             {t_printf*1e6:9.1f} us  {mb/t_printf/1e3:7.3f} GB/s")
        L.nav_host_free(pin)
        ctx.close()


if __name__ == "__main__":
    main()
