"""Runs only the blur measurement hook (for ncu captures): python tools/blur_probe.py [n_img h w iters]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfx_image_stitching_b200 import _capi

n, h, w, iters = (int(x) for x in (sys.argv[1:5] + ['18', '1024', '768', '5'][len(sys.argv) - 1:]))
ctx = _capi.default_context(0)
for s in (1.2262734984654078, 1.5450077936447955, 1.9465878414647133, 2.4525469969308156, 3.0900155872895909):
    ms = C.c_float()
    _capi.check(ctx.lib.b200sift_bench_blur(ctx.handle, n, h, w, s, iters, 1, C.byref(ms)))
    print(f'sigma {s:.4f}: {ms.value * 1e3:.1f} us  {8.0 * n * h * w / ms.value / 1e6:.0f} GB/s')
