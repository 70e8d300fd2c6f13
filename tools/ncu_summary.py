#!/usr/bin/env python
"""Summarise an ncu --page raw --csv export: the handful of metrics the roofline discussion needs.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv && python tools/ncu_summary.py raw.csv
"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.avg.per_cycle_active', 'sm__cycles_elapsed.max']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print('---', d[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                print(f"  {w:70s} {d[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
        vals = []
        for h in stall:
            try:
                vals.append((float(d[hdr.index(h)].replace(',', '')), h))
            except ValueError:
                pass
        for v, h in sorted(vals, reverse=True)[:6]:
            print(f"  stall {v:8.2f}%  {h.split('warp_issue_stalled_')[1].split('_per_warp')[0]}")


if __name__ == "__main__":
    main(sys.argv[1])
