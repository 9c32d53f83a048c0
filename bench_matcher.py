#!/usr/bin/env python
"""Matcher sweep of BASELINE.json configs[4]: all-pairs nearest-neighbour over NA x NB 128-d uint8
descriptors on one B200, tensor-core kernel timed alone (CUDA events), as a fraction of the dense
INT8 tensor peak measured on this pool (profiles/r2_int8_peak.json: cuBLASLt IGEMM 8192^3 through
torch._int_mm; the matcher issues tcgen05.mma kind::i8) and of the measured bf16 peak of
MEASURED_PEAKS.json.  Also checks a sample of the result against the oracle.

    python bench_matcher.py [--sizes 2048 4096 ...] [--out profiles/r1_matcher_sweep.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', type=int, nargs='*', default=[2048, 4096, 8192, 16384, 32768, 65536])
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    from vfx_image_stitching_b200 import _capi
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    ctx = _capi.default_context()
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'bf16_tflops': 1590.0}
    peak = float(peaks['bf16_tflops'])
    i8p = os.path.join(ROOT, 'profiles', 'r2_int8_peak.json')
    peak_i8 = float(json.load(open(i8p))['int8_tops']) if os.path.exists(i8p) else 2.0 * peak
    rows = []
    shapes = [(n, n) for n in a.sizes] + [(a.sizes[0], a.sizes[-1]), (a.sizes[-1], a.sizes[0])]
    for na, nb in shapes:
        for top2 in (0, 1):
            ms = C.c_float()
            _capi.check(ctx.lib.b200sift_bench_match(ctx.handle, None, na, None, nb, top2, a.iters, C.byref(ms)))
            tops = 2.0 * 128 * na * nb / (ms.value * 1e-3) / 1e12
            rows.append({'nA': na, 'nB': nb, 'epilogue': 'top2' if top2 else 'best', 'ms': ms.value,
                         'Tops': tops, 'frac_of_measured_int8_peak': tops / peak_i8,
                         'frac_of_measured_bf16_peak': tops / peak,
                         'desc_pairs_per_s': na * nb / (ms.value * 1e-3)})
            print(json.dumps(rows[-1]), flush=True)
    # correctness spot check through the public API (host in / host out) against the oracle
    from oracle import sift_oracle as so
    rng = np.random.default_rng(8)
    A = rng.integers(0, 256, (3000, 128), dtype=np.uint8)
    B = rng.integers(0, 256, (5000, 128), dtype=np.uint8)
    B[4000] = B[17]
    A[5] = B[17]
    idx, d1, d2 = iss.match_descriptors(A, B, return_second=True)
    r = so.match_u8(A, B)
    ok = bool(np.array_equal(idx, r[0]) and np.array_equal(d1, r[1]) and np.array_equal(d2, r[2]))
    print(json.dumps({'oracle_check_3000x5000': ok}), flush=True)
    if a.out:
        json.dump({'peak_bf16_tflops_measured': peak, 'peak_int8_tops_measured': peak_i8,
                   'note': 'int8 peak: cuBLASLt IGEMM 8192^3 burst (tools/measure_int8_peak.py)',
                   'rows': rows, 'oracle_check': ok}, open(a.out, 'w'), indent=1)
    assert ok


if __name__ == '__main__':
    main()
