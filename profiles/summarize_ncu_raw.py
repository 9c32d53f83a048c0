#!/usr/bin/env python
"""Key metrics per kernel from `ncu ... --csv --page raw` output (live --log-file or `ncu -i rep --page raw --csv`).
usage: python profiles/summarize_ncu_raw.py raw.csv [every_nth_launch]"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size']


def main(path, nth=1):
    lines = [l for l in open(path) if not l.startswith('==') and l.strip()]
    rows = list(csv.reader(lines))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    for n, r in enumerate(data):
        if n % nth != nth - 1:
            continue
        print('==', r[ix['Kernel Name']][:60], 'grid', r[ix['Grid Size']], 'block', r[ix['Block Size']])
        for k in KEYS:
            if k in ix:
                print(f'   {k:75s} {r[ix[k]]:>16s} {units[ix[k]]}')
        st = sorted(((float(r[ix[h]]), h) for h in stall if r[ix[h]]), reverse=True)[:5]
        print('   top stalls (warps per issue): ' + ', '.join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
