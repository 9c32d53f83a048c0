#!/usr/bin/env python
"""Golden fixture for row f2: the UNMODIFIED reference's cylindrical_projection
(/root/reference/image_stitching_sift.py:117-136) on raw inputs.

Test infrastructure only; runs in the build container (needs /root/reference).  Records, in
tests/golden/cyl.npz:
  * the raw (cv2.imread) images of out/ in pano.txt order + their focal lengths + the reference's
    projected outputs (the images the CLI hands to SIFT, image_stitching_sift.py:295);
  * small crops of the first parrington / grail images at their own focal lengths, a short focal
    length (strong curvature: many collisions, pixels falling outside) and a long one, incl. odd
    sizes, each with the reference's output.
tests/test_oracle_golden.py pins oracle.cylindrical_projection to these bit for bit; the GPU
test then compares the CUDA kernels with the same outputs.

Usage: python tests/golden/make_golden_cyl.py
"""
import os
import sys

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    import cv2
    import image_stitching_sift as iss
    out = {}
    paths, focals = iss.read_pano_data(os.path.join(REF, 'out', 'pano.txt'))
    names = []
    for i, (p, f) in enumerate(zip(paths, focals)):
        base = p.replace('\\', '/').split('/')[-1]
        raw = cv2.imread(os.path.join(REF, 'out', base))
        assert raw is not None, base
        names.append(os.path.splitext(base)[0])
        out[f'out_raw_{i}'] = raw
        out[f'out_cyl_{i}'] = iss.cylindrical_projection(raw, f)
    out['out_names'] = np.array(names)
    out['out_focals'] = np.array(focals, np.float64)
    cases = []
    for folder, crop in (('parrington', (slice(100, 261), slice(60, 247))), ('grail', (slice(0, 150), slice(200, 384)))):
        ps, fs = iss.read_pano_data(os.path.join(REF, folder, 'pano.txt'))
        base = ps[0].replace('\\', '/').split('/')[-1]
        raw = cv2.imread(os.path.join(REF, folder, base))
        assert raw is not None, base
        c = np.ascontiguousarray(raw[crop])
        for f in (fs[0], 60.0, 2500.5):
            cases.append((c, float(f)))
    rng = np.random.default_rng(3)
    cases.append((rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), 41.25))
    cases.append((rng.integers(0, 256, (2, 3, 3), dtype=np.uint8), 5.0))
    for k, (img, f) in enumerate(cases):
        out[f'case_img_{k}'] = img
        out[f'case_focal_{k}'] = np.float64(f)
        out[f'case_out_{k}'] = iss.cylindrical_projection(img, f)
    out['n_cases'] = np.int64(len(cases))
    out['meta'] = np.array(f'cv2 {cv2.__version__} numpy {np.__version__}')
    np.savez_compressed(os.path.join(HERE, 'cyl.npz'), **out)
    print('wrote cyl.npz:', len(cases), 'cases +', len(names), 'out images')


if __name__ == '__main__':
    main()
