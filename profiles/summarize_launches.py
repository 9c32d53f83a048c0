#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main(path, steps=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    names = [re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name'])) for r in rows]
    # whole steps only: a step starts with the grey + upsample kernel; whatever follows the last
    # complete step (e.g. the roofline launches bench.py makes after its timed loops) is dropped
    starts = [i for i, k in enumerate(names) if k.endswith('gray_upsample_kernel')]
    if len(starts) >= 2 and steps is None:
        rows, names = rows[starts[0]:starts[-1]], names[starts[0]:starts[-1]]
        steps = len(starts) - 1
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row, k in zip(rows, names):
        v = float(row['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[row['Metric Unit']]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    if steps is None:
        steps = 1
    print(f'{n} launches, {tot:.0f} us total, {steps} step(s) in the capture -> {tot / steps:.0f} us, '
          f'{n / steps:.0f} launches per step')
    print(f'{"us/step":>10} {"n/step":>7} {"us each":>9} {"share":>7}  kernel')
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{t / steps:10.1f} {c / steps:7.1f} {t / c:9.1f} {100 * t / tot:6.1f}%  {k[:100]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
