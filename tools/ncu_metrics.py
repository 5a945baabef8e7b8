#!/usr/bin/env python
"""Print the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` output.
    python tools/ncu_metrics.py raw.csv [extra_metric_substring ...]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fp64.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_bytes.sum']
extra = sys.argv[2:]
for d in data:
    print('----')
    for i, h in enumerate(hdr):
        if h in want or any(e in h for e in extra):
            print('%-75s %s %s' % (h, d[i][:100], units[i]))
