#!/usr/bin/env python
"""Per-LAUNCH list of an .ncu-rep (ncu -i X --page raw --csv): kernel, grid, duration, DRAM bytes read / written,
executed warp instructions, issue-active.  With --json OUT also writes the traffic file bench.py reads
(profiles/rN_traffic.json: last launch of every kernel; frames_per_launch recorded for the sequence kernel).
usage: ncu_kernels.py X.ncu-rep [--json OUT --frames-per-launch K]"""
import csv
import io
import json
import subprocess
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    path = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def unit_scale(col):   # to bytes / microseconds
        u = units[idx[col]]
        return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3, "msecond": 1e3,
                "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
    print(f"# {path}: one line per launch (ncu --set full: caches flushed before every launch, kernels serialised)")
    print(f"{'kernel':28s} {'grid':>8s} {'us':>9s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'warp inst':>11s} {'issue':>6s}")
    last = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("nav::", "")
        us = f(r[idx["gpu__time_duration.sum"]]) * unit_scale("gpu__time_duration.sum")
        rd = f(r[idx["dram__bytes_read.sum"]]) * unit_scale("dram__bytes_read.sum")
        wr = f(r[idx["dram__bytes_write.sum"]]) * unit_scale("dram__bytes_write.sum")
        inst = f(r[idx["smsp__inst_executed.sum"]])
        iss = r[idx["smsp__issue_active.avg.per_cycle_active"]] if "smsp__issue_active.avg.per_cycle_active" in idx else ""
        print(f"{name[:28]:28s} {r[idx['launch__grid_size']]:>8s} {us:9.2f} {rd / 1e6:11.3f} {wr / 1e6:11.3f} {inst:11.0f} {iss:>6s}")
        last[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "us": us, "warp_instructions": inst}
    if "--json" in sys.argv:
        out = sys.argv[sys.argv.index("--json") + 1]
        k = int(sys.argv[sys.argv.index("--frames-per-launch") + 1]) if "--frames-per-launch" in sys.argv else None
        if k:
            for name in last:
                if "k_frame_seq" in name:
                    last[name]["frames_per_launch"] = k
        with open(out, "w") as fo:
            json.dump({"source": f"ncu --set full --clock-control none ({path})", "kernels": last}, fo, indent=1)


if __name__ == "__main__":
    main()
