"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv > raw.csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('---', r[ki][:40])
    for w in want:
        if w in hdr:
            print(f"   {w:72s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
