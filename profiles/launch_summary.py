#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X).
usage: launch_summary.py launches.csv "command that was profiled" """
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    t = tot[r["Kernel Name"]]
    t[0] += us
    t[1] += 1
total = sum(t[0] for t in tot.values())
print(f"# per-kernel totals of {sys.argv[1]} (ncu --metrics gpu__time_duration.sum, cold-cache serialised)")
if len(sys.argv) > 2:
    print(f"# command: {sys.argv[2]}")
for name, (us, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {cnt:5d}x {us / cnt:10.2f} us/launch {100 * us / total:6.1f}%  {name[:90]}")
