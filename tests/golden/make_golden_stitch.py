#!/usr/bin/env python
"""Golden vectors for row f4 (read_pano_data, pad_image, blend_two_images, rectangle_crop, the
drift-corrected second loop of run_panorama), produced by the UNMODIFIED reference.

Test infrastructure only; needs /root/reference, so it runs in the build container.  Writes
tests/golden/stitch.npz:
  * small synthetic blend cases (inputs + the reference's output), all sign combinations of the
    shift, Python-float and numpy-float64 ref_match (numpy blends in float32 resp. float64);
  * rectangle_crop cases;
  * read_pano_data of the three pano.txt files;
  * out/: the full mosaic of the reference's second loop + crop; parrington/, grail/: shape and
    SHA-256 of the mosaic and of the cropped result (the images themselves are too large to commit).
    Inputs are the projected images stored in tests/golden/<set>.npz (colour for out/, the grey
    projections replicated to 3 channels for the 18-image sets).
Shifts and matched pairs come from tests/golden/<set>.npz, i.e. from the reference's own
compute_shift_sift.

Usage: python tests/golden/make_golden_stitch.py
"""
import hashlib
import os
import sys

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import cv2  # noqa: E402
import image_stitching_sift as ref  # noqa: E402  (the unmodified reference)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synthetic_cases():
    rng = np.random.default_rng(42)
    cases = []
    specs = [(30, 40, 30, 36, (17.3, 2.6)), (30, 40, 30, 40, (-21.5, -3.4)), (28, 33, 28, 33, (9.5, -4.5)),
             (24, 50, 24, 30, (-30.2, 0.0)), (20, 20, 20, 20, (19.6, 1.2)), (26, 31, 26, 29, (12.0, 0.5))]
    for k, (ha, wa, hb, wb, (dx, dy)) in enumerate(specs):
        a = rng.integers(0, 256, (ha, wa, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (hb, wb, 3), dtype=np.uint8)
        a[:, :3] = 0          # black borders like a cylindrical projection
        b[:, -4:] = 0
        a[:2] = 0
        if k == 4:
            b[:, 5:9] = 0     # a black column inside the overlap
        xa, ya = wa * 0.7 + 0.37, ha * 0.5
        pair = ((xa, ya), (xa - dx, ya - dy))
        for strong in (False, True):
            rm = tuple(tuple(np.float64(v) if strong else float(v) for v in p) for p in pair)
            out = ref.blend_two_images((dx, dy), rm, a, b)
            cases.append(dict(a=a, b=b, shift=np.array([dx, dy]), pair=np.array(pair).reshape(4), strong=strong, out=out))
    return cases


def crop_cases():
    rng = np.random.default_rng(5)
    cases = []
    for k, (h, w, thr, margin) in enumerate([(40, 60, 0, 3), (40, 60, 12, 15), (33, 21, 0, 0), (16, 16, 0, 2)]):
        img = np.zeros((h, w, 3), np.uint8)
        if k != 3:
            y0, y1, x0, x1 = 5 + k, h - 6, 4, w - 3 - k
            img[y0:y1, x0:x1] = rng.integers(0, 40, (y1 - y0, x1 - x0, 3), dtype=np.uint8)
        out = ref.rectangle_crop(img, thr, margin)
        cases.append(dict(img=img, thr=thr, margin=margin, out=out))
    return cases


def second_loop(name):
    """cylindrical images of the set + golden shifts/pairs -> the reference's mosaic and crop."""
    g = np.load(os.path.join(HERE, name + '.npz'))
    folder = os.path.join(REF, name)
    paths, focals = ref.read_pano_data(os.path.join(folder, 'pano.txt'))
    # inputs that travel with the repo: the colour projections of out/ (fixture bgr_0, bgr_1); for the
    # 18-image sets the fixture's grey projections replicated to 3 channels
    if 'bgr_0' in g.files and all(f'bgr_{i}' in g.files for i in range(len(paths))):
        cyl = [g[f'bgr_{i}'].copy() for i in range(len(paths))]
    else:
        cyl = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g['gray']]
    shift_list = [tuple(float(v) for v in s) for s in g['shifts']]
    pairs = [((float(b[0]), float(b[1])), (float(b[2]), float(b[3]))) for b in g['best_pairs']]
    # image_stitching_sift.py:336-381, literally
    acc_shifts = [(0, 0)]
    for i in range(len(shift_list)):
        prev_x, prev_y = acc_shifts[i]
        cur_dx, cur_dy = shift_list[i]
        acc_shifts.append((prev_x + cur_dx, prev_y + cur_dy))
    final_dx, final_dy = acc_shifts[-1]
    N = len(cyl)
    average_drift = final_dy / (N - 1) if N > 1 else 0
    new_shift_list = [(dx, dy - average_drift) for dx, dy in shift_list]
    mosaic = cyl[0].copy()
    for i in range(1, N):
        diff_y = mosaic.shape[0] - cyl[i].shape[0]
        if diff_y != 0:
            cyl[i] = ref.pad_image(cyl[i], 0, diff_y)
        mosaic = ref.blend_two_images(new_shift_list[i - 1], pairs[i - 1], mosaic, cyl[i])
    crop = ref.rectangle_crop(mosaic, 0, 15)
    return mosaic, crop, paths, focals


def main():
    out = {}
    sc = synthetic_cases()
    out['n_blend'] = len(sc)
    for i, c in enumerate(sc):
        for k, v in c.items():
            out[f'blend{i}_{k}'] = np.asarray(v)
    cc = crop_cases()
    out['n_crop'] = len(cc)
    for i, c in enumerate(cc):
        for k, v in c.items():
            out[f'crop{i}_{k}'] = np.asarray(v)
    for name in ('out', 'parrington', 'grail'):
        mosaic, crop, paths, focals = second_loop(name)
        out[f'{name}_paths'] = np.array(paths)
        out[f'{name}_focals'] = np.array(focals, np.float64)
        out[f'{name}_mosaic_shape'] = np.array(mosaic.shape)
        out[f'{name}_crop_shape'] = np.array(crop.shape)
        out[f'{name}_mosaic_sha'] = np.array(sha(mosaic))
        out[f'{name}_crop_sha'] = np.array(sha(crop))
        if name == 'out':
            out['out_mosaic'] = mosaic
        print(name, mosaic.shape, crop.shape, sha(mosaic)[:12])
    out['meta'] = np.array(f'cv2 {cv2.__version__} numpy {np.__version__}')
    np.savez_compressed(os.path.join(HERE, 'stitch.npz'), **out)
    print('wrote', os.path.join(HERE, 'stitch.npz'), os.path.getsize(os.path.join(HERE, 'stitch.npz')))


if __name__ == '__main__':
    main()
