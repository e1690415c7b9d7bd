#!/usr/bin/env python
"""Hot source lines of a kernel from an .ncu-rep captured with --import-source on:
ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > mix.csv
ncu_lines.py mix.csv [top]   -> per (file, line): share of executed warp instructions and of stall samples"""
import csv
import sys
from collections import defaultdict

cur_file = None
hdr = None
agg = defaultdict(lambda: [0, 0, ""])
line_no, line_src = None, ""
for row in csv.reader(open(sys.argv[1])):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row
        i_addr = hdr.index("Address")
        i_ins = hdr.index("Instructions Executed")
        i_smp = hdr.index("# Samples")
        continue
    if hdr is None:
        continue
    if row[0].strip():          # a CUDA source line
        line_no, line_src = row[0].strip(), row[1]
        continue
    if len(row) > i_ins and row[i_addr].strip():   # a SASS line belonging to the last CUDA line
        k = (cur_file, line_no)
        agg[k][0] += int(row[i_ins]) if row[i_ins].strip().isdigit() else 0
        agg[k][1] += int(row[i_smp]) if row[i_smp].strip().isdigit() else 0
        agg[k][2] = line_src
tot_i = sum(v[0] for v in agg.values())
tot_s = sum(v[1] for v in agg.values())
print(f"{tot_i} warp instructions, {tot_s} samples")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:>4}  instr {100 * v[0] / max(tot_i, 1):5.1f}%  samples {100 * v[1] / max(tot_s, 1):5.1f}%  {v[2].strip()[:100]}")
