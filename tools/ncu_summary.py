#!/usr/bin/env python3
"""Condenses `ncu -i REPORT --page raw --csv` into the per-kernel summary lines kept under profiles/
(kernel | metric | value | unit; the LAST captured launch of each kernel name).
  ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py "header text" > profiles/x_summary.txt"""
import csv, sys
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
kcol = names.index("Kernel Name")
last = {}
order = []
for r in data:
    if len(r) != len(names): continue
    k = r[kcol]
    if k not in last: order.append(k)
    last[k] = r
print("# " + (sys.argv[1] if len(sys.argv) > 1 else "ncu --set full --clock-control none"))
print("# last captured launch of each kernel; kernel | metric | value | unit")
for k in order:
    r = last[k]
    for m in KEEP:
        if m in names:
            i = names.index(m)
            print("%-60s | %s | %s | %s" % (k[:60], m, r[i], units[i]))
