#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline needs.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size',
        'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio','launch__waves_per_multiprocessor',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__cycles_active.avg','smsp__cycles_active.avg']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70], 'id', r[hdr.index('ID')])
    for w in WANT:
        if w in hdr:
            print(f'  {w:80s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
