#!/usr/bin/env python
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Sizes are tiny on purpose (the tools slow kernels down 10-100x): the packed ring blur with partial
strips and batches, the pyramid (ring + tile + tail kernels), extrema / refine / orientation /
descriptor kernels, the per-image sort, the tcgen05 matcher with > 3 B tiles per CTA, finalize + vote,
the stage-API kernels.  Results are compared with the oracle so that a tool-induced change would show.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from conftest import natural_image
    from oracle import sift_oracle as so
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    from vfx_image_stitching_b200 import sift_impl as si
    rng = np.random.default_rng(0)
    img = (rng.random((130, 513)) * 255).astype(np.float32)
    for s in (1.2262735, 1.9465878, 3.0900156):
        assert np.abs(si.gaussian_blur(img, s) - so.gaussian_blur(img, s, 'c')).max() < 2e-4
    a = natural_image(200, 260, 11, channels=3)
    b = np.roll(a, (3, -17), axis=(0, 1))
    res = si.detect_and_describe_batch([a, b])
    shifts, nm, best, bp = iss.match_pairs([(0, 1), (1, 0)])
    print('keypoints', [len(k) for k, _ in res], 'matches', nm.tolist(), 'shifts', shifts)
    A = rng.integers(0, 256, (300, 128), dtype=np.uint8)
    B = rng.integers(0, 256, (1800, 128), dtype=np.uint8)      # 8 B tiles: ring and TMEM buffers wrap
    idx, d1, d2 = iss.match_descriptors(A, B, return_second=True)
    ridx, r1, r2 = so.match_u8(A, B)
    assert np.array_equal(idx, ridx) and np.array_equal(d1, r1) and np.array_equal(d2, r2)
    ia, ib = iss.ratio_test_matches(res[0][1], res[1][1])
    g = so.to_gray_float(a)
    base = so.generate_base_image(g, 1.6, 0.5)
    pyr = so.generate_gaussian_images(base, so.compute_number_of_octaves(base.shape), so.generate_gaussian_kernels(1.6, 3))
    cand = si.extrema_candidates(pyr)
    kps, lyr = si.localize_extrema(cand, pyr)
    raw = si.find_scale_space_extrema_array(pyr)
    d = si.generate_descriptors(raw[:50], pyr, window_width=3, num_bins=6)
    print('candidates', len(cand), 'localized', int((lyr >= 0).sum()), 'oriented', len(raw), 'ratio-test', len(ia), d.shape)
    print('sanitize_run ok')


if __name__ == '__main__':
    main()
