#!/usr/bin/env python
"""Blur kernel alone at the two bench shapes for every radius of the reference's sigma chain (same hook as bench.py)."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfx_image_stitching_b200 import _capi, sift_impl
ctx = _capi.default_context(0)
sig = sift_impl.generate_gaussian_kernels(1.6, 3)
names = [('R5', 1.2489996)] + [(f'l{l}', float(sig[l])) for l in range(1, 6)]
out = {}
for shape, iters in (((18, 1024, 768), 20), ((8, 6144, 8192), 5)):
    row = []
    for name, s in names:
        ms = C.c_float()
        _capi.check(ctx.lib.b200sift_bench_blur(ctx.handle, shape[0], shape[1], shape[2], s, iters, 1, C.byref(ms)))
        row.append(round(ms.value * 1e3, 1))
    by = 8.0 * shape[0] * shape[1] * shape[2]
    mean = sum(row) / len(row)
    out[str(shape)] = {'us': row, 'mean_us': round(mean, 1), 'frac': round(by / (mean * 1e-6) / 6531.9e9, 3)}
print(json.dumps(out))
