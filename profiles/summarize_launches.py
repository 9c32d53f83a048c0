#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main(path, steps=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    first = None
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', row['Kernel Name'])
        k = re.sub(r'^void ', '', k)
        v = float(row['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[row['Metric Unit']]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
        if first is None:
            first = k
    if steps is None:
        steps = max(1, agg.get('b200::gray_upsample_kernel', [1])[0])
    print(f'{n} launches, {tot:.0f} us total, {steps} step(s) in the capture -> {tot / steps:.0f} us, '
          f'{n / steps:.0f} launches per step')
    print(f'{"us/step":>10} {"n/step":>7} {"us each":>9} {"share":>7}  kernel')
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{t / steps:10.1f} {c / steps:7.1f} {t / c:9.1f} {100 * t / tot:6.1f}%  {k[:100]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
