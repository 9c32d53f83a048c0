#!/usr/bin/env python
"""Instruction mix + stall samples per opcode from `ncu --page source --csv` of ONE kernel.
usage: ncu -i rep.ncu-rep --page source --csv -k regex:<kernel> -c 1 [-s N] | python profiles/inst_mix.py"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
col = {h: i for i, h in enumerate(rows[start])}
tot = 0
byop = collections.Counter()
samples = collections.Counter()
for r in rows[start + 1:]:
    if len(r) < 10 or not r[col['Instructions Executed']].isdigit():
        continue
    ins = int(r[col['Instructions Executed']])
    tot += ins
    toks = r[col['Source']].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    byop[op] += ins
    samples[op] += int(r[col['# Samples']])
print(rows[0][1][:100] if rows[0] else '')
print('total warp instructions', tot, ' total stall samples', sum(samples.values()))
for op, c in byop.most_common(30):
    print(f'{op:10s} {c:11d} {100 * c / tot:5.1f}%   samples {samples[op]:7d} {100 * samples[op] / max(1, sum(samples.values())):5.1f}%')
