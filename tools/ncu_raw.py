#!/usr/bin/env python3
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv`. usage: ncu -i rep --page raw --csv | python tools/ncu_raw.py [extra-regex]"""
import csv, re, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'launch__shared_mem_per_block_dynamic']
extra = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
for i, h in enumerate(hdr):
    if h in keys or (extra and extra.search(h)):
        print("%-75s %-12s %s" % (h, units[i], [r[i][:60] for r in data]))
