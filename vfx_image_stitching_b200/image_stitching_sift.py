"""Drop-in for the hot-path functions of the reference's `image_stitching_sift` module.

compute_shift_sift (image_stitching_sift.py:52-83): two detect+describe calls, the brute-force
A->B nearest-neighbour matcher (:63-79) and the ransac() translation vote (:86-111), all on the GPU.
cylindrical_projection (:117-136) is the step before the path.

`panorama_shifts` is the batched form of run_panorama's first loop (:312-327): every image is
detected once (the reference recomputes each interior image twice), adjacent pairs are matched on
the device-resident descriptors.
"""
import ctypes as C
import math

import numpy as np

from . import sift_impl
from ._capi import check, default_context, ptr


# ----------------------------------------------------------------------------- matcher
def match_descriptors(descA, descB, ctx=None, return_second=False):
    """Nearest (and second nearest) row of descB for every row of descA, exact squared L2.

    The arithmetic of image_stitching_sift.py:64-73: strict '<' arg-min, lowest j wins ties.
    Descriptors must be the integer-valued 0..255 vectors generate_descriptors emits (uint8 or
    float32 holding integers).  Returns (best_idx int32, best_d2 int32[, second_d2 int32]).
    """
    ctx = ctx or default_context()
    A = _as_u8(descA)
    B = _as_u8(descB)
    idx = np.full(len(A), -1, np.int32)
    d1 = np.full(len(A), np.iinfo(np.int32).max, np.int32)
    d2 = np.full(len(A), np.iinfo(np.int32).max, np.int32)
    if len(A):
        check(ctx.lib.b200sift_match(ctx.handle, ptr(A), len(A), ptr(B) if len(B) else None, len(B), 0, ptr(idx),
                                     ptr(d1), ptr(d2)))
    return (idx, d1, d2) if return_second else (idx, d1)


def _as_u8(desc):
    d = np.asarray(desc)
    if d.size == 0:
        return np.zeros((0, 128), np.uint8)
    d = d.reshape(-1, 128)
    if d.dtype != np.uint8:
        r = np.rint(d)
        if not (np.array_equal(r, d) and r.min() >= 0 and r.max() <= 255):
            raise ValueError('descriptors must be integer valued in 0..255 (what generate_descriptors returns)')
        d = r.astype(np.uint8)
    return np.ascontiguousarray(d)


def _int_thresh(desc_thresh):
    """The reference compares an integer-valued distance with `desc_thresh` (strict <, :74): for any
    real threshold t that is d < ceil(t)."""
    return int(min(max(math.ceil(desc_thresh), -(2 ** 31) + 1), 2 ** 31 - 1))


def ratio_test_matches(descA, descB, ratio=0.7, ctx=None, return_distances=False):
    """The good-match filter of sift_visualizeUI.py:247-257 (knnMatch(k=2), m.distance < 0.7 * n.distance)
    with EXACT nearest / second-nearest neighbours and exact integer arithmetic: row i of descA is kept iff
    den^2 * d1 < num^2 * d2 on the squared distances, ratio = num / den (7 / 10 for the reference's 0.7;
    other ratios are taken as the closest fraction with denominator <= 1000).  The reference asks
    approximate FLANN KD-trees for the neighbours, so its own list is not reproducible run to run.
    Returns (ia, ib) int32 -- query rows in order and their nearest train rows -- and, when asked,
    the squared distances (d1, d2) of every query row."""
    from fractions import Fraction
    ctx = ctx or default_context()
    fr = Fraction(ratio).limit_denominator(1000)
    if fr <= 0:
        raise ValueError('ratio must be positive')
    A = _as_u8(descA)
    B = _as_u8(descB)
    ia = np.zeros(len(A), np.int32)
    ib = np.zeros(len(A), np.int32)
    d1 = np.full(len(A), np.iinfo(np.int32).max, np.int32)
    d2 = np.full(len(A), np.iinfo(np.int32).max, np.int32)
    n = C.c_int32(0)
    if len(A):
        check(ctx.lib.b200sift_ratio_match(ctx.handle, ptr(A), len(A), ptr(B) if len(B) else None, len(B), 0,
                                           fr.numerator, fr.denominator, ptr(ia), ptr(ib), ptr(d1), ptr(d2),
                                           C.byref(n)))
    ia, ib = ia[:n.value].copy(), ib[:n.value].copy()
    return (ia, ib, d1, d2) if return_distances else (ia, ib)


def good_matches(descA, descB, ratio=0.7, ctx=None):
    """ratio_test_matches as the list of cv2.DMatch the GUI draws (sift_visualizeUI.py:254-257):
    queryIdx / trainIdx / distance = Euclidean distance of the pair."""
    import cv2
    ia, ib, d1, _ = ratio_test_matches(descA, descB, ratio, ctx, return_distances=True)
    return [cv2.DMatch(int(i), int(j), float(np.sqrt(np.float32(d1[i])))) for i, j in zip(ia, ib)]


def match_keypoints(kpsA, descA, kpsB, descB, desc_thresh=25000, ctx=None):
    """The match list of image_stitching_sift.py:63-79 -> (ia, ib, [((xA,yA),(xB,yB)), ...])."""
    idx, d2 = match_descriptors(descA, descB, ctx)
    keep = (d2 < _int_thresh(desc_thresh)) & (idx != -1)
    ia = np.nonzero(keep)[0].astype(np.int32)
    ib = idx[keep]
    matches = [(kpsA[i].pt, kpsB[j].pt) for i, j in zip(ia, ib)]
    return ia, ib, matches


# ----------------------------------------------------------------------------- ransac (f1)
def ransac(matches, dist_sq_thresh=3, ctx=None):
    """image_stitching_sift.py:86-111 -> (best_move (dx, dy), best_pair or None); first maximum wins."""
    if len(matches) == 0:
        return (0, 0), None
    ctx = ctx or default_context()
    m = np.ascontiguousarray(np.asarray(matches, dtype=np.float64).reshape(-1, 4))   # Python floats, as :94-96
    mv = (C.c_double * 2)()
    best = C.c_int32()
    check(ctx.lib.b200sift_ransac(ctx.handle, ptr(m), len(m), float(dist_sq_thresh), mv, C.byref(best)))
    return (mv[0], mv[1]), matches[best.value]


def compute_shift_sift(imgA, imgB, ransac_thr=3, desc_thresh=25000):
    """image_stitching_sift.py:52-83 -> (best_move, best_pair)."""
    ctx = default_context()
    a = np.asarray(imgA)
    b = np.asarray(imgB)
    if a.shape == b.shape and a.dtype == b.dtype:
        sift_impl.detect_and_describe_batch([a, b], ctx=ctx, download=False)
        shifts, nm, _, bp = match_pairs([(0, 1)], ransac_thr, desc_thresh, ctx)
        return shifts[0], bp[0]
    else:  # different shapes: two passes, host descriptors
        kA, dA = sift_impl.compute_keypoints_and_descriptors(a)
        kB, dB = sift_impl.compute_keypoints_and_descriptors(b)
        _, _, matches = match_keypoints(kA, dA, kB, dB, desc_thresh, ctx)
    return ransac(matches, dist_sq_thresh=ransac_thr, ctx=ctx)


# ----------------------------------------------------------------------------- projection (f2)
def cylindrical_projection(img_bgr, focal_len, ctx=None):
    """image_stitching_sift.py:117-136 (forward map, last writer wins) on the GPU."""
    ctx = ctx or default_context()
    img = np.ascontiguousarray(img_bgr, np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    check(ctx.lib.b200sift_cylindrical_projection(ctx.handle, ptr(img), h, w, ch, float(focal_len), ptr(out)))
    return out


# ----------------------------------------------------------------------------- batched first loop
def panorama_shifts(images, ransac_thr=3, desc_thresh=25000, ctx=None, return_details=False):
    """Shifts of all adjacent pairs (i, i+1) of a list of same-shape images: the loop of
    image_stitching_sift.py:312-327 with each image detected once.  Returns [(dx, dy), ...]
    (and per-pair details when asked)."""
    ctx = ctx or default_context()
    counts = sift_impl.detect_and_describe_batch(images, ctx=ctx, download=False)
    pairs = [(i, i + 1) for i in range(len(images) - 1)]
    shifts, n_matches, best, _ = match_pairs(pairs, ransac_thr, desc_thresh, ctx)
    if not return_details:
        return shifts
    details = []
    for p, n in enumerate(n_matches):
        ia = np.zeros(int(n), np.int32)
        ib = np.zeros(int(n), np.int32)
        if n:
            check(ctx.lib.b200sift_get_pair_matches(ctx.handle, p, ptr(ia), ptr(ib), None))
        details.append(dict(n_matches=int(n), ia=ia, ib=ib, best=int(best[p])))
    return shifts, counts, details


def match_pairs(pairs, ransac_thr=3, desc_thresh=25000, ctx=None):
    """Matcher + acceptance + vote for a list of (imgA, imgB) index pairs of the last
    detect_and_describe_batch, in one device pass.  Returns (shifts [(dx, dy)], n_matches,
    best_index, best_pairs [((xA,yA),(xB,yB)) or None])."""
    ctx = ctx or default_context()
    n = len(pairs)
    if n == 0:
        return [], np.zeros(0, np.int32), np.zeros(0, np.int32), []
    pr = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
    sh = np.zeros((n, 2), np.float64)
    nm = np.zeros(n, np.int32)
    best = np.zeros(n, np.int32)
    xy = np.zeros((n, 4), np.float32)
    check(ctx.lib.b200sift_match_pairs(ctx.handle, n, pr.ctypes.data_as(C.POINTER(C.c_int32)), _int_thresh(desc_thresh),
                                       float(ransac_thr), sh.ctypes.data_as(C.POINTER(C.c_double)),
                                       nm.ctypes.data_as(C.POINTER(C.c_int32)),
                                       best.ctypes.data_as(C.POINTER(C.c_int32)), ptr(xy)))
    shifts = [(float(sh[p, 0]), float(sh[p, 1])) if nm[p] else (0, 0) for p in range(n)]
    bp = [((float(xy[p, 0]), float(xy[p, 1])), (float(xy[p, 2]), float(xy[p, 3]))) if nm[p] else None
          for p in range(n)]
    return shifts, nm, best, bp


# ----------------------------------------------------------------------------- f4: after the path
def read_pano_data(pano_file_path):
    """image_stitching_sift.py:12-46 -> (image paths, focal lengths) of an AutoStitch pano.txt.

    Host logic: a line naming a .jpg / .png opens an entry, the next blank-free line that parses
    as a float closes it; everything else (sizes, matrices, blank lines) is skipped."""
    images, focuses = [], []
    pending = None
    with open(pano_file_path, 'r', encoding='utf-8') as fh:
        for raw in fh.read().splitlines():
            token = raw.strip()
            low = token.lower()
            if '.jpg' in low or '.png' in low:
                pending = token
                continue
            if not low or ' ' in low:
                continue
            try:
                focal = float(low)
            except ValueError:
                continue
            if pending is not None:
                images.append(pending)
                focuses.append(focal)
                pending = None
    return images, focuses


def pad_image(img_bgr, move_x, move_y):
    """image_stitching_sift.py:139-153: translate by zero padding (rounded to whole pixels)."""
    mx, my = int(round(move_x)), int(round(move_y))
    img = np.asarray(img_bgr)
    h, w = img.shape[:2]
    out = np.zeros((h + abs(my), w + abs(mx), 3), img.dtype)
    out[max(my, 0):max(my, 0) + h, max(mx, 0):max(mx, 0) + w] = img
    return out


def blend_two_images(shift_vec, ref_match, imgA, imgB, ctx=None):
    """image_stitching_sift.py:156-202 on the GPU: imgB joined to imgA with a column-wise linear
    cross-fade over the overlap.  Bit-identical to the reference, including numpy's choice of
    float32 arithmetic when ref_match holds Python floats (the CLI) and float64 when it holds numpy
    float64 scalars."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(imgA, np.uint8)
    b = np.ascontiguousarray(imgB, np.uint8)
    if a.ndim != 3 or b.ndim != 3 or a.shape[2] != 3 or b.shape[2] != 3:
        raise ValueError('blend_two_images needs two H x W x 3 images')
    strong = int(any(isinstance(v, np.floating) for p in ref_match for v in p))
    rm = (C.c_double * 4)(float(ref_match[0][0]), float(ref_match[0][1]), float(ref_match[1][0]),
                          float(ref_match[1][1]))
    oh, ow = C.c_int32(), C.c_int32()
    args = (ctx.handle, ptr(a), a.shape[0], a.shape[1], ptr(b), b.shape[0], b.shape[1], float(shift_vec[0]),
            float(shift_vec[1]), rm, strong)
    check(ctx.lib.b200sift_blend_two_images(*args, None, 0, C.byref(oh), C.byref(ow)))
    out = np.empty((oh.value, ow.value, 3), np.uint8)
    check(ctx.lib.b200sift_blend_two_images(*args, ptr(out), out.nbytes, C.byref(oh), C.byref(ow)))
    return out


def rectangle_crop(img, black_threshold, extra_margin, ctx=None):
    """image_stitching_sift.py:208-247: crop to the bounding box of the pixels brighter than
    black_threshold (grey value of cv2.cvtColor), trimmed by extra_margin at the top and bottom."""
    ctx = ctx or default_context()
    im = np.asarray(img)
    src = np.ascontiguousarray(im, np.uint8)
    h, w = src.shape[:2]
    box = (C.c_int32 * 4)()
    check(ctx.lib.b200sift_crop_bbox(ctx.handle, ptr(src), h, w, int(black_threshold), box))
    y_min, y_max, x_min, x_max = (int(v) for v in box)
    if y_max < 0:
        return img
    y_min = max(0, y_min + extra_margin)
    y_max = min(h - 1, y_max - extra_margin)
    if y_min > y_max or x_min > x_max:
        return img
    return im[y_min:y_max + 1, x_min:x_max + 1]


def drift_corrected_shifts(shift_list, n_images):
    """image_stitching_sift.py:336-365: remove the accumulated vertical drift, spread evenly."""
    total_dy = 0
    for _, dy in shift_list:
        total_dy = total_dy + dy
    average_drift = total_dy / (n_images - 1) if n_images > 1 else 0
    return [(dx, dy - average_drift) for dx, dy in shift_list]


def stitch_panorama(cyl_imgs, ransac_thr=3, desc_thresh=25000, margin=15, ctx=None):
    """Both loops of run_panorama (image_stitching_sift.py:312-384) on already projected images:
    adjacent-pair shifts (every image detected once, pairs matched on the device), drift
    correction, chained blends, crop.  Returns (result, mosaic, shift_list, matched_pairs)."""
    ctx = ctx or default_context()
    imgs = [np.asarray(im) for im in cyl_imgs]
    for i in range(len(imgs) - 1):                      # :318-320 equalise heights pairwise
        diff_y = imgs[i].shape[0] - imgs[i + 1].shape[0]
        if diff_y != 0:
            imgs[i + 1] = pad_image(imgs[i + 1], 0, diff_y)
    same = all(im.shape == imgs[0].shape for im in imgs)
    if same and len(imgs) > 1:
        sift_impl.detect_and_describe_batch(imgs, ctx=ctx, download=False)
        shift_list, _, _, matched_pairs = match_pairs([(i, i + 1) for i in range(len(imgs) - 1)], ransac_thr,
                                                      desc_thresh, ctx)
    else:
        shift_list, matched_pairs = [], []
        for i in range(len(imgs) - 1):
            s, p = compute_shift_sift(imgs[i], imgs[i + 1], ransac_thr, desc_thresh)
            shift_list.append(s)
            matched_pairs.append(p)
    new_shifts = drift_corrected_shifts(shift_list, len(imgs))
    mosaic = imgs[0].copy()
    for i in range(1, len(imgs)):
        nxt = imgs[i]
        diff_y = mosaic.shape[0] - nxt.shape[0]
        if diff_y != 0:
            nxt = pad_image(nxt, 0, diff_y)
        mosaic = blend_two_images(new_shifts[i - 1], matched_pairs[i - 1], mosaic, nxt, ctx=ctx)
    return rectangle_crop(mosaic, 0, margin, ctx=ctx), mosaic, shift_list, matched_pairs


def run_panorama(folder_path=None, pano_file=None, margin=None, save=True):
    """image_stitching_sift.py:253-389.  With no arguments it asks the same three questions as the
    reference; passing them makes the run non-interactive.  Returns the cropped panorama (and writes
    <folder>/panoroma_sift.jpg like the reference when `save`)."""
    import os
    import cv2
    if folder_path is None:
        folder_path = input('image folder (default .): ').strip()
    folder_path = folder_path or '.'
    if not folder_path.endswith(('/', '\\')):
        folder_path += '/'
    if pano_file is None:
        pano_file = input('pano.txt path (default <folder>/pano.txt): ').strip()
    pano_file = pano_file or folder_path + 'pano.txt'
    img_paths, focals = read_pano_data(pano_file)
    if not img_paths:
        print('no usable entries in', pano_file)
        return None
    cyl = []
    for p, f in zip(img_paths, focals):
        full = p if os.path.exists(p) else os.path.join(folder_path, os.path.basename(p.replace('\\', '/')))
        img = cv2.imread(full)
        if img is None:
            raise FileNotFoundError(full)
        cyl.append(cylindrical_projection(img, f))
    if margin is None:
        m = input('crop margin (default 15): ').strip()
        margin = int(m) if m.isdigit() else 15
    result, _, _, _ = stitch_panorama(cyl, margin=int(margin))
    if save:
        cv2.imwrite(os.path.join(folder_path, 'panoroma_sift.jpg'), result)
    return result
