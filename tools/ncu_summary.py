#!/usr/bin/env python
"""Key metrics of every launch in an .ncu-rep (reads `ncu -i rep --page raw --csv`).  usage: ncu_summary.py rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum']
for r in rows[2:]:
    print("--- launch")
    for k in want:
        if k in hdr:
            i = hdr.index(k); print(f"{k} [{units[i]}] = {r[i]}")
