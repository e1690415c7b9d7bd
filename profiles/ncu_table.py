#!/usr/bin/env python
"""Compact per-kernel table from an .ncu-rep (reads it with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
]


def main(path, which="last"):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    byk = {}
    for r in rows[2:]:
        byk.setdefault(r[idx['Kernel Name']][:30], []).append(r)
    names = list(byk)
    print(f"{'metric':72s}" + "".join(f"{n[:17]:>18s}" for n in names))
    for w in WANT:
        if w not in idx:
            continue
        line = f"{w[:72]:72s}"
        for n in names:
            rs = byk[n]
            r = rs[-1] if which == "last" else max(rs, key=lambda x: float(x[idx['gpu__time_duration.sum']].replace(',', '')))
            line += f"{r[idx[w]][:16]:>18s}"
        print(line + "  " + units[idx[w]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "last")
