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
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        # L2 traffic of the tex/LSU unit split by the eviction class of the request: the TMA-staged gt / logits stream
        # is issued with an L2 evict-first policy, every other read of the kernel (packed template taps, summed-area
        # table, theta, meshgrid factors — >95 % of them template taps and table entries) is evict-normal.  So
        # evict_normal hit / (hit + miss) IS the L2 hit rate on the template traffic (north star: "L2 hit rate on the
        # template"), and the l1tex global-load lookups are its L1 hit rate (the gt stream bypasses L1 through TMA).
        "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"]
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

    def ratio(hit, miss):
        if hit in hdr and miss in hdr:
            out_ = []
            for r in rows[2:]:
                try:
                    h_, m_ = float(r[hdr.index(hit)]), float(r[hdr.index(miss)])
                    out_.append("%.4f" % (h_ / (h_ + m_)) if h_ + m_ > 0 else "")
                except ValueError:
                    out_.append("")
            return out_
        return None
    t = ratio("lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum",
              "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum")
    if t:
        w.writerow(["derived: L2 hit rate on template / table reads (evict-normal tex reads)", "fraction"] + t)
    t = ratio("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
              "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum")
    if t:
        w.writerow(["derived: L1 hit rate on template / table reads (LSU global loads)", "fraction"] + t)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
open("/tmp/_sass.csv", "w").write(sass)
import os
hist = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_sass_hist.py"), "/tmp/_sass.csv", "0"],
                      capture_output=True, text=True).stdout
open(out + "_sass_hist.txt", "w").write(hist)
print("wrote", out + "_metrics.csv", out + "_sass_hist.txt")
