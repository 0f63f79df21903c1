#!/usr/bin/env python
"""Write a small, committable summary of an .ncu-rep: key metrics per profiled launch (CSV) plus
the opcode mix / stall breakdown of the first kernel.  Usage: ncu_summary.py rep.ncu-rep out_prefix"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__inst_executed_pipe_tma.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out + "_metrics.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
open("/tmp/_sass.csv", "w").write(sass)
import os
hist = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_sass_hist.py"), "/tmp/_sass.csv", "0"],
                      capture_output=True, text=True).stdout
open(out + "_sass_hist.txt", "w").write(hist)
print("wrote", out + "_metrics.csv", out + "_sass_hist.txt")
