#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep (raw page), as a compact table / csv.

    python tools/ncu_summary.py rep.ncu-rep [out.csv]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [hdr.index("Kernel Name")] + [hdr.index(k) for k in KEYS if k in hdr]
out = [[hdr[i] for i in cols], [units[i] for i in cols]] + [[r[i] for i in cols] for r in rows[2:]]
if len(sys.argv) > 2:
    with open(sys.argv[2], "w", newline="") as fh:
        csv.writer(fh).writerows(out)
for r in rows[2:]:
    print("----", r[cols[0]][:60])
    for i in cols[1:]:
        print(f"  {hdr[i]:<66} {r[i]:>16} {units[i]}")
