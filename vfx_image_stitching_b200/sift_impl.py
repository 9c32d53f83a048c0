"""Drop-in for the reference's `sift_impl` module, running on one B200.

Same function names, positional order, keyword defaults and return types as
/root/reference/sift_impl.py:15-526 (numpy images in; list of cv2.KeyPoint and
float32 (N,128) descriptors out), so that

    from vfx_image_stitching_b200 import sift_impl            # instead of `import sift_impl`
    from vfx_image_stitching_b200.sift_impl import compute_keypoints_and_descriptors

works for image_stitching_sift.py:6 and sift_visualizeUI.py:16,104-115.  All
arithmetic of the path runs in hand-written sm_100a kernels behind the C ABI in
include/b200sift.h; this file only converts between numpy / cv2 objects and the
flat buffers of that ABI.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import KP_DTYPE, B200SiftError, check, default_context, default_params, ptr, ptr_array

try:  # the reference returns genuine cv2.KeyPoint objects (FLANN / drawing need them)
    import cv2 as _cv2
    _KeyPoint = _cv2.KeyPoint
except Exception:  # pragma: no cover - cv2 is part of the reference's requirements
    _cv2 = None

    class _KeyPoint:  # minimal stand-in with the fields the reference uses
        def __init__(self, x=0.0, y=0.0, size=0.0, angle=-1.0, response=0.0, octave=0, class_id=-1):
            self.pt = (float(np.float32(x)), float(np.float32(y)))
            self.size = float(np.float32(size))
            self.angle = float(np.float32(angle))
            self.response = float(np.float32(response))
            self.octave = int(octave)
            self.class_id = int(class_id)

# 全局微小數值容差 of the reference (sift_impl.py:9)
float_tolerance = 1e-7


# ----------------------------------------------------------------------------- helpers
def keypoints_to_array(keypoints):
    """list[cv2.KeyPoint] -> structured array (x, y, size, angle, response, octave)."""
    out = np.zeros(len(keypoints), KP_DTYPE)
    for i, k in enumerate(keypoints):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave)
    return out


def array_to_keypoints(arr):
    """structured array -> list[cv2.KeyPoint] (class_id = -1 as in the reference)."""
    return [_KeyPoint(float(r['x']), float(r['y']), float(r['size']), float(r['angle']), float(r['response']),
                      int(r['octave'])) for r in arr]


def _as_pyramid(gaussian_images):
    """Sequence [octave][layer] of 2-D arrays -> (flat list of C-contiguous float32 arrays, h, w, n_oct, n_layers)."""
    n_oct = len(gaussian_images)
    if n_oct == 0:
        raise ValueError('empty pyramid')
    n_layers = len(gaussian_images[0])
    flat = []
    for o in range(n_oct):
        if len(gaussian_images[o]) != n_layers:
            raise ValueError('ragged pyramid')
        for l in range(n_layers):
            flat.append(np.ascontiguousarray(gaussian_images[o][l], dtype=np.float32))
    h, w = flat[0].shape
    for o in range(n_oct):
        eh, ew = h >> o, w >> o
        for l in range(n_layers):
            if flat[o * n_layers + l].shape != (eh, ew):
                raise ValueError(f'octave {o} layer {l} has shape {flat[o * n_layers + l].shape}, expected {(eh, ew)}')
    return flat, h, w, n_oct, n_layers


def _object_pyramid(layers, n_oct, n_layers):
    out = np.empty((n_oct, n_layers), dtype=object)
    for o in range(n_oct):
        for l in range(n_layers):
            out[o, l] = layers[o * n_layers + l]
    return out


# ----------------------------------------------------------------------------- batch entry points
def detect_and_describe_batch(images, sigma=1.6, num_intervals=3, assumed_blur=0.5, image_border_width=5,
                              ctx=None, download=True):
    """compute_keypoints_and_descriptors for a list of same-shape images in one device pass.

    Returns a list of (keypoint structured array, uint8 (N,128) descriptors) per image when
    `download`, else only the per-image counts (results stay on the device for match_images /
    the multi-GPU all-gather).
    """
    ctx = ctx or default_context()
    lib = ctx.lib
    if len(images) and hasattr(images[0], 'data_ptr'):
        return _detect_device_tensors(images, sigma, num_intervals, assumed_blur, image_border_width, ctx, download)
    imgs = [np.asarray(im) for im in images]
    if not imgs:
        return []
    first = imgs[0]
    if any(im.shape != first.shape for im in imgs):
        raise ValueError('detect_and_describe_batch needs images of identical shape')
    if first.ndim == 3 and first.shape[2] == 3:
        if first.dtype != np.uint8:
            raise TypeError('3-channel input must be uint8 BGR (cv2.imread layout), as in the reference CLI')
        channels, dtype = 3, 0
        imgs = [np.ascontiguousarray(im, np.uint8) for im in imgs]
    elif first.ndim == 2:
        channels = 1
        if first.dtype == np.uint8:
            dtype = 0
            imgs = [np.ascontiguousarray(im) for im in imgs]
        else:  # image.astype('float32') of the reference (sift_impl.py:29)
            dtype = 1
            imgs = [np.ascontiguousarray(im, np.float32) for im in imgs]
    else:
        raise ValueError(f'unsupported image shape {first.shape}')
    h, w = first.shape[:2]
    p = default_params(sigma=sigma, num_intervals=num_intervals, assumed_blur=assumed_blur,
                       image_border_width=image_border_width)
    counts = np.zeros(len(imgs), np.int32)
    check(lib.b200sift_detect_describe(ctx.handle, C.byref(p), len(imgs), ptr_array(imgs), h, w, channels, dtype,
                                       imgs[0].strides[0], 0, counts.ctypes.data_as(C.POINTER(C.c_int32))))
    if not download:
        return counts
    return download_results(counts, ctx)


def download_results(counts, ctx=None, out=None):
    """(keypoints, uint8 descriptors) of every image of the last detect, copied to the host in two
    transfers.  `counts` must be the counts of ALL images of that detect (the C side refuses a
    short buffer).  `out` = optional (kps, desc) host arrays to fill (e.g. views of pinned memory)."""
    ctx = ctx or default_context()
    counts = np.asarray(counts, np.int64)
    total = int(counts.sum())
    if out is None:
        kps = np.empty(total, KP_DTYPE)
        desc = np.empty((total, 128), np.uint8)
    else:
        kps, desc = out[0][:total], out[1][:total]
    if total:
        check(ctx.lib.b200sift_get_all_keypoints(ctx.handle, ptr(kps), ptr(desc), total))
    off = np.concatenate([[0], np.cumsum(counts)])
    return [(kps[off[i]:off[i + 1]], desc[off[i]:off[i + 1]]) for i in range(len(counts))]


def _detect_device_tensors(images, sigma, num_intervals, assumed_blur, image_border_width, ctx, download):
    """Same as detect_and_describe_batch for images that already live in HBM (torch CUDA tensors,
    uint8 HxWx3 BGR / HxW, or float32 HxW; C-contiguous rows)."""
    first = images[0]
    shape = tuple(first.shape)
    if any(tuple(t.shape) != shape for t in images):
        raise ValueError('detect_and_describe_batch needs images of identical shape')
    channels = 3 if len(shape) == 3 else 1
    if channels == 3 and shape[2] != 3:
        raise ValueError(f'unsupported image shape {shape}')
    esz = first.element_size()
    if esz not in (1, 4) or (esz == 4 and channels != 1):
        raise TypeError('device images must be uint8 (BGR or grey) or float32 grey')
    dtype = 0 if esz == 1 else 1
    for t in images:
        if not t.is_cuda or t.stride(-1) != 1:
            raise ValueError('device images must be CUDA tensors with contiguous pixels')
    h, w = shape[:2]
    arr = (C.c_void_p * len(images))(*[t.data_ptr() for t in images])
    p = default_params(sigma=sigma, num_intervals=num_intervals, assumed_blur=assumed_blur,
                       image_border_width=image_border_width)
    counts = np.zeros(len(images), np.int32)
    check(ctx.lib.b200sift_detect_describe(ctx.handle, C.byref(p), len(images), arr, h, w, channels, dtype,
                                           int(first.stride(0)) * esz, 1, counts.ctypes.data_as(C.POINTER(C.c_int32))))
    return download_results(counts, ctx) if download else counts


def stage_stats(image_index=0, ctx=None):
    """(#3x3x3 extrema, #localized, #oriented) of the last detect for one image."""
    ctx = ctx or default_context()
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    check(ctx.lib.b200sift_get_stats(ctx.handle, int(image_index), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


# ----------------------------------------------------------------------------- reference API
def compute_keypoints_and_descriptors(image, sigma=1.6, num_intervals=3, assumed_blur=0.5, image_border_width=5):
    """sift_impl.py:15-39 -> (list[cv2.KeyPoint], float32 ndarray (N,128); shape (0,) when N == 0)."""
    (kps, desc), = detect_and_describe_batch([image], sigma, num_intervals, assumed_blur, image_border_width)
    if len(kps) == 0:
        return [], np.array([], dtype='float32')  # sift_impl.py:526 on an empty list
    return array_to_keypoints(kps), desc.astype(np.float32)


def generate_base_image(image, sigma, assumed_blur):
    """sift_impl.py:45-56: 2x bilinear upsample + Gaussian blur sqrt(sigma^2 - (2*assumed_blur)^2)."""
    ctx = default_context()
    if sigma is None:
        sigma = 1.6
    img = np.ascontiguousarray(image, np.float32)
    if img.ndim != 2:
        raise ValueError('generate_base_image expects a 2-D grey image')
    out = np.empty((2 * img.shape[0], 2 * img.shape[1]), np.float32)
    check(ctx.lib.b200sift_base_image(ctx.handle, ptr(img), img.shape[0], img.shape[1], float(sigma),
                                      float(assumed_blur), ptr(out)))
    return out


def compute_number_of_octaves(image_shape):
    """sift_impl.py:59-63"""
    return int(np.round(np.log(min(image_shape)) / np.log(2) - 1))


def generate_gaussian_kernels(sigma, num_intervals):
    """sift_impl.py:66-79 (pure parameter arithmetic; the device derives the same values from `sigma`)."""
    num_images_per_octave = num_intervals + 3
    k = 2 ** (1. / num_intervals)
    kernels = np.zeros(num_images_per_octave)
    kernels[0] = sigma
    for idx in range(1, num_images_per_octave):
        sigma_prev = (k ** (idx - 1)) * sigma
        sigma_total = k * sigma_prev
        kernels[idx] = np.sqrt(sigma_total ** 2 - sigma_prev ** 2)
    return kernels


def gaussian_blur(image, sigma):
    """cv2.GaussianBlur(image, (0, 0), sigmaX=sigma, sigmaY=sigma) on float32 (sift_impl.py:56,91)."""
    ctx = default_context()
    img = np.ascontiguousarray(image, np.float32)
    out = np.empty_like(img)
    check(ctx.lib.b200sift_gaussian_blur(ctx.handle, ptr(img), img.shape[0], img.shape[1], float(sigma), ptr(out), 0))
    return out


def generate_gaussian_images(image, num_octaves, gaussian_kernels):
    """sift_impl.py:82-97 -> object ndarray [num_octaves][len(kernels)] of float32 images."""
    ctx = default_context()
    base = np.ascontiguousarray(image, np.float32)
    sig = np.ascontiguousarray(gaussian_kernels, np.float64)
    n_layers = len(sig)
    h, w = base.shape
    layers = [np.empty((h >> o, w >> o), np.float32) for o in range(num_octaves) for _ in range(n_layers)]
    check(ctx.lib.b200sift_gaussian_pyramid(ctx.handle, ptr(base), h, w, int(num_octaves),
                                            sig.ctypes.data_as(C.POINTER(C.c_double)), n_layers, ptr_array(layers)))
    return _object_pyramid(layers, num_octaves, n_layers)


def generate_DoG_images(gaussian_images):
    """sift_impl.py:100-111 -> object ndarray [n_oct][n_layers-1]."""
    ctx = default_context()
    flat, h, w, n_oct, n_layers = _as_pyramid(gaussian_images)
    dogs = [np.empty((h >> o, w >> o), np.float32) for o in range(n_oct) for _ in range(n_layers - 1)]
    check(ctx.lib.b200sift_dog_pyramid(ctx.handle, ptr_array(flat), h, w, n_oct, n_layers, ptr_array(dogs)))
    return _object_pyramid(dogs, n_oct, n_layers - 1)


def find_scale_space_extrema(gaussian_images, dog_images, num_intervals, sigma, border, contrast_threshold=0.04):
    """sift_impl.py:117-140 -> list[cv2.KeyPoint] in the reference's scan order (base-image coordinates).

    Extrema and the quadratic fit read `dog_images` (:124-129), orientations read `gaussian_images`
    (:135-136), like the reference.  With dog_images=None the DoG is the float32 difference of
    adjacent Gaussian layers (what generate_DoG_images returns), formed on the fly and never stored.
    """
    arr = find_scale_space_extrema_array(gaussian_images, num_intervals, sigma, border, contrast_threshold,
                                         dog_images=dog_images)
    return array_to_keypoints(arr)


def find_scale_space_extrema_array(gaussian_images, num_intervals=3, sigma=1.6, border=5, contrast_threshold=0.04,
                                   dog_images=None):
    ctx = default_context()
    flat, h, w, n_oct, n_layers = _as_pyramid(gaussian_images)
    dflat, dptr = None, None
    if dog_images is not None:
        dflat, dh, dw, d_oct, d_layers = _as_pyramid(dog_images)
        if (dh, dw, d_oct, d_layers) != (h, w, n_oct, n_layers - 1):
            raise ValueError('dog_images does not match gaussian_images')
        dptr = ptr_array(dflat)
    p = default_params(sigma=sigma, num_intervals=num_intervals, image_border_width=border,
                       contrast_threshold=contrast_threshold)
    cap = 1 << 15
    while True:
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int32()
        rc = ctx.lib.b200sift_find_extrema(ctx.handle, C.byref(p), ptr_array(flat), dptr, h, w, n_oct, n_layers,
                                           ptr(out), cap, C.byref(n))
        if rc == -3 and n.value > cap:  # B200SIFT_ECAPACITY: the caller's buffer was too small
            cap = int(n.value)
            continue
        check(rc)
        return out[:n.value].copy()


def extrema_candidates(gaussian_images, num_intervals=3, border=5, contrast_threshold=0.04):
    """All (octave, layer, y, x) that pass is_pixel_an_extremum (sift_impl.py:143-163), in scan order."""
    ctx = default_context()
    flat, h, w, n_oct, n_layers = _as_pyramid(gaussian_images)
    p = default_params(num_intervals=num_intervals, image_border_width=border, contrast_threshold=contrast_threshold)
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 4), np.int32)
        n = C.c_int32()
        rc = ctx.lib.b200sift_extrema_candidates(ctx.handle, C.byref(p), ptr_array(flat), h, w, n_oct, n_layers,
                                                 ptr(out), cap, C.byref(n))
        if rc == -3 and n.value > cap:
            cap = int(n.value)
            continue
        check(rc)
        return out[:n.value].copy()


def is_pixel_an_extremum(prev_patch, curr_patch, next_patch, threshold):
    """sift_impl.py:143-163 on three 3x3 views (27 comparisons; the device scan is extrema_candidates)."""
    val = curr_patch[1, 1]
    if abs(val) <= threshold:
        return False
    cube = np.stack([prev_patch, curr_patch, next_patch])
    return bool(np.all(val >= cube)) if val > 0 else bool(np.all(val <= cube))


def localize_extrema(cands, layers, num_intervals=3, sigma=1.6, contrast_threshold=0.04, border=5, eigen_ratio=10,
                     max_iter=5, is_dog=False, octave_base=-1):
    """Batched localize_extremum_via_quadratic_fit (sift_impl.py:169-211): `cands` = (n, 4) int32 rows
    (octave, layer, y, x) -- the format extrema_candidates returns; `layers` = a [octave][layer]
    pyramid of Gaussian layers (is_dog=False) or DoG layers (is_dog=True), or, with octave_base >= 0,
    the single octave `octave_base` as [[layer, ...]].  Returns (keypoint structured array (n),
    final_layer int32 (n), -1 where the reference returns None)."""
    ctx = default_context()
    cands = np.ascontiguousarray(np.asarray(cands, np.int32).reshape(-1, 4))
    flat, h, w, n_oct, n_layers = _as_pyramid(layers)
    p = default_params(sigma=sigma, num_intervals=num_intervals, image_border_width=border,
                       contrast_threshold=contrast_threshold, eigen_ratio=eigen_ratio, max_iter=max_iter)
    kps = np.zeros(len(cands), KP_DTYPE)
    lyr = np.full(len(cands), -1, np.int32)
    check(ctx.lib.b200sift_localize(ctx.handle, C.byref(p), ptr_array(flat), int(bool(is_dog)), h, w, n_oct, n_layers,
                                    int(octave_base), ptr(cands), len(cands), ptr(kps), ptr(lyr)))
    return kps, lyr


def localize_extremum_via_quadratic_fit(x, y, layer, octave, num_intervals, dog_octave, sigma, contrast_threshold,
                                        border, eigen_ratio=10, max_iter=5):
    """sift_impl.py:169-211 -> (cv2.KeyPoint, final layer) or None.  `dog_octave` = dog_images[octave]
    (the num_intervals + 2 DoG layers of that octave), exactly what the reference is called with (:128-129)."""
    kps, lyr = localize_extrema([(octave, layer, y, x)], [list(dog_octave)], num_intervals, sigma, contrast_threshold,
                                border, eigen_ratio, max_iter, is_dog=True, octave_base=octave)
    if lyr[0] < 0:
        return None
    k = kps[0]
    kp = _KeyPoint()
    kp.pt = (float(k['x']), float(k['y']))
    kp.octave = int(k['octave'])
    kp.size = float(k['size'])
    kp.response = float(k['response'])
    return kp, int(lyr[0])


def keypoints_with_orientations(kps, octave, gauss_img, radius_factor=3, num_bins=36, peak_ratio=0.8,
                                scale_factor=1.5):
    """Batched compute_keypoints_with_orientations (sift_impl.py:246-293) for a keypoint structured array
    on one Gaussian image: returns (oriented keypoints (m) in input order / ascending bin, counts (n))."""
    ctx = default_context()
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    img = np.ascontiguousarray(gauss_img, np.float32)
    p = default_params(radius_factor=radius_factor, ori_bins=num_bins, peak_ratio=peak_ratio, scale_factor=scale_factor)
    out = np.zeros((len(kps), int(num_bins)), KP_DTYPE)
    counts = np.zeros(len(kps), np.int32)
    check(ctx.lib.b200sift_orientations(ctx.handle, C.byref(p), ptr(kps), len(kps), int(octave), ptr(img), img.shape[0],
                                        img.shape[1], ptr(out), ptr(counts)))
    flat = np.concatenate([out[i, :counts[i]] for i in range(len(kps))]) if len(kps) else out.reshape(-1)
    return flat, counts


def compute_keypoints_with_orientations(keypoint, octave, gauss_img, radius_factor=3, num_bins=36, peak_ratio=0.8,
                                        scale_factor=1.5):
    """sift_impl.py:246-293 -> list[cv2.KeyPoint], one per histogram peak (ascending bin)."""
    arr = keypoints_to_array([keypoint])
    out, _ = keypoints_with_orientations(arr, octave, gauss_img, radius_factor, num_bins, peak_ratio, scale_factor)
    return array_to_keypoints(out)


def compute_gradient_at_center_pixel(cube):
    """sift_impl.py:217-224"""
    dx = 0.5 * (cube[1, 1, 2] - cube[1, 1, 0])
    dy = 0.5 * (cube[1, 2, 1] - cube[1, 0, 1])
    ds = 0.5 * (cube[2, 1, 1] - cube[0, 1, 1])
    return np.array([dx, dy, ds])


def compute_hessian_at_center_pixel(cube):
    """sift_impl.py:227-240"""
    v = cube[1, 1, 1]
    dxx = cube[1, 1, 2] - 2 * v + cube[1, 1, 0]
    dyy = cube[1, 2, 1] - 2 * v + cube[1, 0, 1]
    dss = cube[2, 1, 1] - 2 * v + cube[0, 1, 1]
    dxy = 0.25 * (cube[1, 2, 2] - cube[1, 2, 0] - cube[1, 0, 2] + cube[1, 0, 0])
    dxs = 0.25 * (cube[2, 1, 2] - cube[2, 1, 0] - cube[0, 1, 2] + cube[0, 1, 0])
    dys = 0.25 * (cube[2, 2, 1] - cube[2, 0, 1] - cube[0, 2, 1] + cube[0, 0, 1])
    return np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]])


def compare_keypoints(kp1, kp2):
    """sift_impl.py:299-311 (host comparator; the device sort uses the same key order)."""
    if kp1.pt[0] != kp2.pt[0]:
        return kp1.pt[0] - kp2.pt[0]
    if kp1.pt[1] != kp2.pt[1]:
        return kp1.pt[1] - kp2.pt[1]
    if kp1.size != kp2.size:
        return kp2.size - kp1.size
    if kp1.angle != kp2.angle:
        return kp1.angle - kp2.angle
    if kp1.response != kp2.response:
        return kp2.response - kp1.response
    return kp2.class_id - kp1.class_id


def remove_duplicate_keypoints(keypoints):
    """sift_impl.py:314-327: sort by compare_keypoints, drop repeats of (pt, size, angle)."""
    if len(keypoints) < 2:
        return keypoints
    ctx = default_context()
    arr = keypoints_to_array(keypoints)
    n = C.c_int32()
    check(ctx.lib.b200sift_remove_duplicates(ctx.handle, ptr(arr), len(arr), C.byref(n)))
    return array_to_keypoints(arr[:n.value])


def convert_keypoints_to_input_image_size(keypoints):
    """sift_impl.py:333-343 (in place on the KeyPoint objects, like the reference)."""
    out = []
    for kp in keypoints:
        kp.pt = (kp.pt[0] * 0.5, kp.pt[1] * 0.5)
        kp.size *= 0.5
        kp.octave = (kp.octave & ~255) | ((kp.octave - 1) & 255)
        out.append(kp)
    return out


def unpack_octave(keypoint):
    """sift_impl.py:349-358"""
    octave = keypoint.octave & 255
    layer = (keypoint.octave >> 8) & 255
    if octave >= 128:
        octave |= -128
    scale = 1 / np.float32(1 << octave) if octave >= 0 else np.float32(1 << -octave)
    return octave, layer, scale


def generate_descriptors(keypoints, gaussian_images, window_width=4, num_bins=8, scale_multiplier=3,
                         descriptor_max_value=0.2):
    """sift_impl.py:361-526 -> float32 (N,128), integer valued; shape (0,) for an empty list."""
    if len(keypoints) == 0:
        return np.array([], dtype='float32')
    if window_width < 1 or num_bins < 1 or window_width * window_width * num_bins > 1024:
        raise ValueError('window_width**2 * num_bins must be in 1..1024')
    ctx = default_context()
    flat, h, w, n_oct, n_layers = _as_pyramid(gaussian_images)
    arr = keypoints if isinstance(keypoints, np.ndarray) else keypoints_to_array(keypoints)
    arr = np.ascontiguousarray(arr, KP_DTYPE)
    p = default_params(scale_multiplier=scale_multiplier, descriptor_max_value=descriptor_max_value,
                       num_intervals=n_layers - 3, window_width=int(window_width), desc_bins=int(num_bins))
    out = np.empty((len(arr), window_width * window_width * num_bins), np.float32)
    check(ctx.lib.b200sift_descriptors(ctx.handle, C.byref(p), ptr(arr), len(arr), ptr_array(flat), h, w, n_oct,
                                       n_layers, ptr(out)))
    return out


__all__ = [
    'float_tolerance', 'compute_keypoints_and_descriptors', 'generate_base_image', 'compute_number_of_octaves',
    'generate_gaussian_kernels', 'generate_gaussian_images', 'generate_DoG_images', 'find_scale_space_extrema',
    'is_pixel_an_extremum', 'localize_extremum_via_quadratic_fit', 'compute_keypoints_with_orientations',
    'localize_extrema', 'keypoints_with_orientations', 'compute_gradient_at_center_pixel', 'compute_hessian_at_center_pixel',
    'compare_keypoints', 'remove_duplicate_keypoints', 'convert_keypoints_to_input_image_size', 'unpack_octave',
    'generate_descriptors', 'detect_and_describe_batch', 'gaussian_blur', 'extrema_candidates', 'stage_stats',
    'keypoints_to_array', 'array_to_keypoints', 'B200SiftError',
]
