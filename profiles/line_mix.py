#!/usr/bin/env python
"""Per-source-line instruction counts + stall samples of ONE kernel.

Joins `ncu --page source --csv` (per SASS address) with the line table of the cubin
(`nvdisasm -g`), because the CSV export of ncu's CUDA-source view carries no metrics.

    cuobjdump -xelf all libb200sift.so                      # -> detect.sm_100a.cubin ...
    ncu -i rep.ncu-rep --page source --csv -k regex:describe_kernel -c 1 > k.csv
    python profiles/line_mix.py k.csv detect.sm_100a.cubin describe_kernel [top_n]
"""
import collections
import csv
import re
import subprocess
import sys

csv_path, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
col = {h: i for i, h in enumerate(rows[start])}
recs = []
for r in rows[start + 1:]:
    if len(r) < 10 or not r[col['Instructions Executed']].isdigit():
        continue
    recs.append((int(r[col['Address']], 16), int(r[col['Instructions Executed']]), int(r[col['# Samples']]),
                 r[col['Source']].strip()))
base = recs[0][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur = None
infn = False
for ln in dis:
    if ln.startswith('.text.') or re.match(r'^\s*\.section\s+\.text\.', ln):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = 'inlined' in ln
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
by = collections.defaultdict(lambda: [0, 0])
ti = ts = 0
for addr, ins, smp, _ in recs:
    k = line_of.get(addr - base, ('?', 0))
    by[k][0] += ins
    by[k][1] += smp
    ti += ins
    ts += smp
print(f'{kname}: warp instructions {ti}, stall samples {ts}')
srcs = {}
for (f, l), (ins, smp) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ''
    if f not in srcs:
        try:
            srcs[f] = open('vfx_image_stitching_b200/csrc/' + f).read().splitlines()
        except OSError:
            srcs[f] = []
    if 0 < l <= len(srcs[f]):
        text = srcs[f][l - 1].strip()[:90]
    print(f'{100 * ins / ti:5.1f}% inst {100 * smp / max(ts, 1):5.1f}% stall  {f}:{l:<4d} {text}')
