"""CPU timing of the caller-side data formats: this library's CSV reader / writer
against the reference's own reader (oracle/_ref) and a printf restatement of its writer (oracle).
Host code only -- no GPU involved.  Usage: python profiles/prof_io.py [frames]"""
import importlib
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle_lib import Oracle, RefLib, ref_available  # noqa: E402

nav = importlib.import_module("nav-slam_b200")
synth = importlib.import_module("nav-slam_b200.synth")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows, cols = 16, 1800
    seq = synth.l9_sequence(n)
    text = synth.l9_csv(seq)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "parsed_data.csv")
        with open(path, "w") as f:
            f.write(text)
        mb = len(text) / 1e6
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            frames, ts = nav.l9_csv_read(path, rows, cols, n)
            best = min(best, time.perf_counter() - t0)
        print(f"reader  {rows}x{cols} x {n} frames, {mb:.1f} MB of CSV")
        print(f"  nav_l9_csv_read        {best*1e3:8.1f} ms  {mb/best:7.1f} MB/s  {n/best:7.1f} frames/s")
        if ref_available(f"{rows}x{cols}"):
            ref = RefLib(rows, cols)
            rbest = min(ref.l9_csv_read(path, n)[2] for _ in range(3))
            rframes = ref.l9_csv_read(path, n)[0]
            assert rframes.tobytes() == frames.tobytes()
            print(f"  L9_LidarProcessData    {rbest*1e3:8.1f} ms  {mb/rbest:7.1f} MB/s  {n/rbest:7.1f} frames/s"
                  f"   (reference, same bytes out)  -> {rbest/best:.1f}x")
    orc = Oracle()
    for (r, c) in ((16, 1800), (64, 2048)):
        g = np.random.default_rng(1).standard_normal((r, c, 3)) * 5000.0
        lp = np.array([12.5, -3.25, 100.0, 0.1, 0.2, 0.3])
        import ctypes as C
        L = nav.load_library()
        cap = r * c * 700
        buf = C.create_string_buffer(cap)       # allocated and touched outside the timed region
        lpp = nav.binding._pos_array(lp)
        best = obest = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            m = L.nav_csv_format_frame(buf, cap, 7, r, c, g.ctypes.data, None, None, lpp, lpp)
            best = min(best, time.perf_counter() - t0)
            t0 = time.perf_counter()
            b = orc.csv_format_frame(7, g, lp, ekf_pos=lp)
            obest = min(obest, time.perf_counter() - t0)
        a = buf.raw[:m]
        assert a == b
        mb = len(a) / 1e6
        print(f"writer  {r}x{c}: {mb:.1f} MB of CSV per frame")
        print(f"  nav_csv_format_frame   {best*1e3:8.1f} ms  {mb/best:7.1f} MB/s  {1/best:7.1f} frames/s")
        print(f"  snprintf (main.c:324)  {obest*1e3:8.1f} ms  {mb/obest:7.1f} MB/s  {1/obest:7.1f} frames/s"
              f"   (same bytes out)  -> {obest/best:.1f}x")


if __name__ == "__main__":
    main()
