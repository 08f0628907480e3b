#!/usr/bin/env python
"""Prints the handful of raw ncu metrics the design notes quote, from `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__thread_inst_executed_per_inst_executed.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__inst_executed_op_branch.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    print('---', r[H.index('Kernel Name')][:60])
    for w in want:
        if w in H:
            print("  %-75s %20s %s" % (w, r[H.index(w)], U[H.index(w)]))
    for i, h in enumerate(H):
        if 'issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                if float(r[i]) > 2:
                    print("  %-75s %20s" % (h.replace('smsp__average_warp_latency_', '').replace('smsp__average_warps_', ''), r[i]))
            except ValueError:
                pass
