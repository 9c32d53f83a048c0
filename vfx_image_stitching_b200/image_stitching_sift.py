"""Drop-in for the hot-path functions of the reference's `image_stitching_sift` module.

compute_shift_sift (image_stitching_sift.py:52-83): two detect+describe calls, the brute-force
A->B nearest-neighbour matcher (:63-79) and the ransac() translation vote (:86-111), all on the GPU.
cylindrical_projection (:117-136) is the step before the path.

`panorama_shifts` is the batched form of run_panorama's first loop (:312-327): every image is
detected once (the reference recomputes each interior image twice), adjacent pairs are matched on
the device-resident descriptors.
"""
import ctypes as C

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


def match_keypoints(kpsA, descA, kpsB, descB, desc_thresh=25000, ctx=None):
    """The match list of image_stitching_sift.py:63-79 -> (ia, ib, [((xA,yA),(xB,yB)), ...])."""
    idx, d2 = match_descriptors(descA, descB, ctx)
    keep = (d2 < desc_thresh) & (idx != -1)
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
    m = np.ascontiguousarray(np.asarray(matches, dtype=np.float64).reshape(-1, 4).astype(np.float32))
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
    check(ctx.lib.b200sift_match_pairs(ctx.handle, n, pr.ctypes.data_as(C.POINTER(C.c_int32)), int(desc_thresh),
                                       float(ransac_thr), sh.ctypes.data_as(C.POINTER(C.c_double)),
                                       nm.ctypes.data_as(C.POINTER(C.c_int32)),
                                       best.ctypes.data_as(C.POINTER(C.c_int32)), ptr(xy)))
    shifts = [(float(sh[p, 0]), float(sh[p, 1])) if nm[p] else (0, 0) for p in range(n)]
    bp = [((float(xy[p, 0]), float(xy[p, 1])), (float(xy[p, 2]), float(xy[p, 3]))) if nm[p] else None
          for p in range(n)]
    return shifts, nm, best, bp
