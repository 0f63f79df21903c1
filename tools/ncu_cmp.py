#!/usr/bin/env python
"""Side-by-side of the headline metrics of ncu reports: ncu_cmp.py a.ncu-rep b.ncu-rep ..."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_global_st.sum']
STALL = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
for k in ('barrier', 'long_scoreboard', 'short_scoreboard', 'wait', 'no_instruction', 'lg_throttle', 'mio_throttle',
          'math_pipe_throttle', 'branch_resolving', 'not_selected', 'membar', 'drain', 'dispatch_stall', 'sleeping', 'imc_miss'):
    WANT.append(STALL % k)
cols = []
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, r = rows[0], rows[-1]
    cols.append({c: r[i] for i, c in enumerate(h)})
for w in WANT:
    vals = [c.get(w, '-') for c in cols]
    print(f"{w[-70:]:70s}", *[f"{v:>16s}" for v in vals])
