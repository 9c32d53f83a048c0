"""ctypes front-end of the CPU oracle (oracle/sift_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.

The functions keep the reference's names and argument order
(/root/reference/sift_impl.py:15-526, image_stitching_sift.py:52-111) but
return plain numpy data: keypoints are a structured array with the fields of
cv2.KeyPoint that the reference uses (x, y, size, angle, response, octave).

`blur='c'` uses the C restatement of cv2.GaussianBlur; `blur='cv2'` calls
cv2.GaussianBlur itself (third-party, same call the reference makes at
sift_impl.py:56,91) so that everything downstream of the pyramid can be
compared with the reference without the blur's last-bit differences.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([('x', np.float32), ('y', np.float32), ('size', np.float32),
                     ('angle', np.float32), ('response', np.float32), ('octave', np.int32)])

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build():
    """Compile oracle/libsift_oracle.so with gcc (oracle/Makefile)."""
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libsift_oracle.so')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, 'sift_oracle.c')):
            build()
        L = C.CDLL(path)
        L.orc_gaussian_ksize.argtypes = [C.c_double]
        L.orc_gaussian_kernel.argtypes = [C.c_int, C.c_double, _fp]
        L.orc_gaussian_blur.argtypes = [_fp, C.c_int, C.c_int, C.c_double, _fp]
        L.orc_num_octaves.argtypes = [C.c_int, C.c_int]
        L.orc_ransac.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.orc_cylindrical_projection.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.orc_gaussian_sigmas.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(_fp)


# ---------------------------------------------------------------- third-party semantics
def bgr2gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) on uint8 (sift_impl.py:27-28)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().orc_bgr2gray(img.ctypes.data_as(C.c_void_p), h, w, img.strides[0], out.ctypes.data_as(C.c_void_p))
    return out


def resize2x(img):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((2 * h, 2 * w), np.float32)
    lib().orc_resize2x_linear(_f(img), h, w, _f(out))
    return out


def gaussian_ksize(sigma):
    return lib().orc_gaussian_ksize(float(sigma))


def gaussian_kernel(sigma):
    ks = gaussian_ksize(sigma)
    t = np.empty(ks, np.float32)
    lib().orc_gaussian_kernel(ks, float(sigma), _f(t))
    return t


def gaussian_blur(img, sigma, blur='c'):
    img = np.ascontiguousarray(img, np.float32)
    if blur == 'cv2':
        import cv2
        return cv2.GaussianBlur(img, (0, 0), sigmaX=sigma, sigmaY=sigma)
    out = np.empty_like(img)
    lib().orc_gaussian_blur(_f(img), img.shape[0], img.shape[1], float(sigma), _f(out))
    return out


def decimate(img):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((h // 2, w // 2), np.float32)
    lib().orc_decimate(_f(img), h, w, _f(out))
    return out


# ---------------------------------------------------------------- sift_impl API
def generate_base_image(image, sigma, assumed_blur, blur='c'):
    """sift_impl.py:45-56"""
    if sigma is None:
        sigma = 1.6
    image = resize2x(image)
    sigma_diff = np.sqrt(max((sigma ** 2) - ((2 * assumed_blur) ** 2), 0.01))
    return gaussian_blur(image, float(sigma_diff), blur)


def compute_number_of_octaves(image_shape):
    """sift_impl.py:59-63"""
    return lib().orc_num_octaves(int(image_shape[0]), int(image_shape[1]))


def generate_gaussian_kernels(sigma, num_intervals):
    """sift_impl.py:66-79"""
    out = np.zeros(num_intervals + 3)
    lib().orc_gaussian_sigmas(float(sigma), int(num_intervals), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def generate_gaussian_images(image, num_octaves, gaussian_kernels, blur='c'):
    """sift_impl.py:82-97 -> list of lists of float32 arrays"""
    pyr = []
    for _ in range(num_octaves):
        octave = [image]
        for g in gaussian_kernels[1:]:
            image = gaussian_blur(image, float(g), blur)
            octave.append(image)
        pyr.append(octave)
        image = decimate(octave[-3])
    return pyr


def generate_DoG_images(gaussian_images):
    """sift_impl.py:100-111"""
    return [[b - a for a, b in zip(o, o[1:])] for o in gaussian_images]


class _Pyr:
    """Flat pointer view of a Gaussian pyramid for the C entry points."""

    def __init__(self, gaussian_images):
        self.keep = [[np.ascontiguousarray(l, np.float32) for l in o] for o in gaussian_images]
        self.n_oct = len(self.keep)
        self.n_layers = len(self.keep[0]) if self.n_oct else 0
        flat = [l for o in self.keep for l in o]
        self.ptrs = (_fp * max(1, len(flat)))(*[_f(l) for l in flat])
        self.hs = (C.c_int * max(1, self.n_oct))(*[o[0].shape[0] for o in self.keep])
        self.ws = (C.c_int * max(1, self.n_oct))(*[o[0].shape[1] for o in self.keep])


def find_scale_space_extrema(gaussian_images, dog_images, num_intervals, sigma, border,
                             contrast_threshold=0.04, return_stats=False):
    """sift_impl.py:117-140 (dog_images is recomputed from gaussian_images: same float32 subtraction)."""
    P = _Pyr(gaussian_images)
    cap = 1 << 16
    while True:
        out = np.zeros(cap, KP_DTYPE)
        cand = np.zeros((cap * 4, 4), np.int32)
        stats = (C.c_int * 2)()
        n = lib().orc_find_scale_space_extrema(P.ptrs, P.hs, P.ws, P.n_oct, int(num_intervals),
                                               C.c_double(sigma), int(border), C.c_double(contrast_threshold),
                                               out.ctypes.data_as(C.c_void_p), cap,
                                               cand.ctypes.data_as(_ip), cap * 4, stats)
        if n >= 0 and stats[0] <= cap * 4:
            break
        cap *= 4
    if return_stats:
        return out[:n].copy(), dict(n_candidates=stats[0], n_localized=stats[1], candidates=cand[:stats[0]].copy())
    return out[:n].copy()


def localize_extremum_via_quadratic_fit(x, y, layer, octave, num_intervals, gaussian_images, sigma,
                                        contrast_threshold, border, eigen_ratio=10, max_iter=5):
    """sift_impl.py:169-211 (takes the Gaussian pyramid; DoG = float32 difference of its layers)."""
    P = _Pyr(gaussian_images)
    kp = np.zeros(1, KP_DTYPE)
    lyr = C.c_int()
    ok = lib().orc_localize(P.ptrs, P.hs, P.ws, P.n_oct, P.n_layers, int(x), int(y), int(layer), int(octave),
                            int(num_intervals), C.c_double(sigma), C.c_double(contrast_threshold), int(border),
                            C.c_double(eigen_ratio), int(max_iter), kp.ctypes.data_as(C.c_void_p), C.byref(lyr))
    return (kp[0], lyr.value) if ok else None


def compute_keypoints_with_orientations(keypoint, octave, gauss_img, radius_factor=3, num_bins=36,
                                        peak_ratio=0.8, scale_factor=1.5):
    """sift_impl.py:246-293"""
    img = np.ascontiguousarray(gauss_img, np.float32)
    kin = np.zeros(1, KP_DTYPE)
    kin[0] = keypoint
    out = np.zeros(64, KP_DTYPE)
    n = lib().orc_orientations(kin.ctypes.data_as(C.c_void_p), int(octave), _f(img), img.shape[0], img.shape[1],
                               C.c_double(radius_factor), int(num_bins), C.c_double(peak_ratio),
                               C.c_double(scale_factor), out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def remove_duplicate_keypoints(kps):
    """sift_impl.py:314-327"""
    kps = np.ascontiguousarray(kps, KP_DTYPE).copy()
    n = lib().orc_remove_duplicates(kps.ctypes.data_as(C.c_void_p), len(kps))
    return kps[:n].copy()


def convert_keypoints_to_input_image_size(kps):
    """sift_impl.py:333-343"""
    kps = np.ascontiguousarray(kps, KP_DTYPE).copy()
    lib().orc_convert_to_input_size(kps.ctypes.data_as(C.c_void_p), len(kps))
    return kps


def generate_descriptors(kps, gaussian_images, window_width=4, num_bins=8, scale_multiplier=3,
                         descriptor_max_value=0.2):
    """sift_impl.py:361-526"""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    if len(kps) == 0:
        return np.array([], dtype='float32')
    P = _Pyr(gaussian_images)
    out = np.zeros((len(kps), window_width * window_width * num_bins), np.float32)
    lib().orc_descriptors(kps.ctypes.data_as(C.c_void_p), len(kps), P.ptrs, P.hs, P.ws, P.n_oct, P.n_layers,
                          int(window_width), int(num_bins), C.c_double(scale_multiplier),
                          C.c_double(descriptor_max_value), _f(out))
    return out


def to_gray_float(image):
    """sift_impl.py:27-29"""
    image = np.asarray(image)
    if image.ndim == 3 and image.shape[2] == 3:
        image = bgr2gray(image)
    return np.ascontiguousarray(image.astype('float32'))


def compute_keypoints_and_descriptors(image, sigma=1.6, num_intervals=3, assumed_blur=0.5,
                                      image_border_width=5, blur='c', return_stats=False):
    """sift_impl.py:15-39"""
    image = to_gray_float(image)
    base = generate_base_image(image, sigma, assumed_blur, blur)
    n_oct = compute_number_of_octaves(base.shape)
    sig = generate_gaussian_kernels(sigma, num_intervals)
    g = generate_gaussian_images(base, n_oct, sig, blur)
    kps, stats = find_scale_space_extrema(g, None, num_intervals, sigma, image_border_width, return_stats=True)
    kps = remove_duplicate_keypoints(kps)
    kps = convert_keypoints_to_input_image_size(kps)
    desc = generate_descriptors(kps, g)
    if return_stats:
        return kps, desc, stats
    return kps, desc


# ---------------------------------------------------------------- matcher / vote / projection
def match_f32(descA, descB):
    """image_stitching_sift.py:63-73 literal float32 loop -> (best_idx, best_d2)"""
    A = np.ascontiguousarray(descA, np.float32).reshape(-1, 128)
    B = np.ascontiguousarray(descB, np.float32).reshape(-1, 128)
    idx = np.full(len(A), -1, np.int32)
    d2 = np.full(len(A), np.inf, np.float32)
    lib().orc_match_f32(_f(A), len(A), _f(B), len(B), 128, idx.ctypes.data_as(_ip), _f(d2))
    return idx, d2


def match_u8(descA, descB):
    """exact integer nearest / second nearest -> (best_idx, best_d2, second_d2)"""
    A = np.ascontiguousarray(descA, np.uint8).reshape(-1, 128)
    B = np.ascontiguousarray(descB, np.uint8).reshape(-1, 128)
    idx = np.full(len(A), -1, np.int32)
    d1 = np.zeros(len(A), np.int32)
    d2 = np.zeros(len(A), np.int32)
    lib().orc_match_u8(A.ctypes.data_as(C.c_void_p), len(A), B.ctypes.data_as(C.c_void_p), len(B), 128,
                       idx.ctypes.data_as(_ip), d1.ctypes.data_as(_ip), d2.ctypes.data_as(_ip))
    return idx, d1, d2


def match_pairs(kpsA, descA, kpsB, descB, desc_thresh=25000):
    """image_stitching_sift.py:63-79 -> (ia, ib, matches n x 4 float64)"""
    idx, d2 = match_f32(descA, descB)
    keep = (d2 < desc_thresh) & (idx != -1)
    ia = np.nonzero(keep)[0].astype(np.int32)
    ib = idx[keep]
    m = np.stack([kpsA['x'][ia], kpsA['y'][ia], kpsB['x'][ib], kpsB['y'][ib]], axis=1).astype(np.float64) \
        if len(ia) else np.zeros((0, 4))
    return ia, ib, m


def ransac(matches, dist_sq_thresh=3):
    """image_stitching_sift.py:86-111 -> ((dx, dy), best_pair or None)"""
    m = np.ascontiguousarray(matches, np.float64).reshape(-1, 4)
    if len(m) == 0:
        return (0, 0), None
    mv = (C.c_double * 2)()
    i = lib().orc_ransac(m.ctypes.data_as(C.POINTER(C.c_double)), len(m), float(dist_sq_thresh), mv)
    return (mv[0], mv[1]), ((m[i, 0], m[i, 1]), (m[i, 2], m[i, 3]))


def compute_shift_sift(imgA, imgB, ransac_thr=3, desc_thresh=25000, blur='c'):
    """image_stitching_sift.py:52-83"""
    ka, da = compute_keypoints_and_descriptors(imgA, blur=blur)
    kb, db = compute_keypoints_and_descriptors(imgB, blur=blur)
    _, _, m = match_pairs(ka, da, kb, db, desc_thresh)
    return ransac(m, ransac_thr)


def cylindrical_projection(img_bgr, focal_len):
    """image_stitching_sift.py:117-136"""
    img = np.ascontiguousarray(img_bgr, np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().orc_cylindrical_projection(img.ctypes.data_as(C.c_void_p), h, w, ch, float(focal_len),
                                     out.ctypes.data_as(C.c_void_p))
    return out


# ----------------------------------------------------------------------------- f4: the step after the path
def read_pano_data(pano_file_path):
    """image_stitching_sift.py:12-46: a line naming a .jpg/.png, followed (not necessarily directly)
    by a line without blanks that parses as float, yields (path, focal)."""
    images, focuses, pending = [], [], None
    with open(pano_file_path, 'r', encoding='utf-8') as f:
        lines = f.read().splitlines()
    for text in lines:
        low = text.strip().lower()
        if '.jpg' in low or '.png' in low:
            pending = text.strip()
        elif ' ' not in low and len(low) > 0:
            try:
                val = float(low)
            except ValueError:
                continue
            if pending is not None:
                images.append(pending)
                focuses.append(val)
                pending = None
    return images, focuses


def pad_image(img_bgr, move_x, move_y):
    """image_stitching_sift.py:139-153: zero padding in front for a non-negative move, behind otherwise."""
    mx, my = int(round(move_x)), int(round(move_y))
    h, w = img_bgr.shape[:2]
    out = np.zeros((h + abs(my), w + abs(mx), 3), img_bgr.dtype)
    oy, ox = max(my, 0), max(mx, 0)
    out[oy:oy + h, ox:ox + w] = img_bgr
    return out


def blend_two_images(shift_vec, ref_match, imgA, imgB):
    """image_stitching_sift.py:156-202, vectorised over columns.  The dtype of alpha follows numpy's
    promotion: a Python float (the reference CLI: ref_match holds kp.pt tuples) blends in float32,
    a numpy float64 scalar in float64."""
    dx, dy = shift_vec
    if dx < 0:
        dx, dy = -dx, -dy
        ref_match = (ref_match[1], ref_match[0])
        imgA, imgB = imgB, imgA
    padA_x = imgB.shape[1] - imgA.shape[1] + ref_match[0][0] - ref_match[1][0]
    padB_x = ref_match[0][0] - ref_match[1][0]
    overlap_range = ref_match[1][0] - ref_match[0][0] + imgA.shape[1]
    shiftA = pad_image(imgA, -padA_x, -dy)
    shiftB = pad_image(imgB, padB_x, dy)
    HH, WW = max(shiftA.shape[0], shiftB.shape[0]), max(shiftA.shape[1], shiftB.shape[1])
    canvasA = np.zeros((HH, WW, 3), np.float32)
    canvasB = np.zeros((HH, WW, 3), np.float32)
    canvasA[:shiftA.shape[0], :shiftA.shape[1]] = shiftA
    canvasB[:shiftB.shape[0], :shiftB.shape[1]] = shiftB
    anyA = (canvasA != 0).any(axis=(0, 2))
    anyB = (canvasB != 0).any(axis=(0, 2))
    both = anyA & anyB
    counter = np.cumsum(both) - both                      # overlap_counter before column cc
    strong = isinstance(overlap_range, np.floating)       # numpy scalar -> float64 arithmetic
    result = np.zeros((HH, WW, 3), np.float32)
    onlyA, onlyB = anyA & ~anyB, anyB & ~anyA
    result[:, onlyA] = canvasA[:, onlyA]
    result[:, onlyB] = canvasB[:, onlyB]
    if both.any():
        alpha = (counter[both] / float(overlap_range)) if overlap_range != 0 else np.zeros(int(both.sum()))
        if strong:
            a64 = alpha.astype(np.float64)[None, :, None]
            result[:, both] = ((1 - a64) * canvasA[:, both].astype(np.float64)
                               + a64 * canvasB[:, both].astype(np.float64)).astype(np.float32)
        else:
            wa = (1.0 - alpha).astype(np.float32)[None, :, None]
            wb = alpha.astype(np.float32)[None, :, None]
            result[:, both] = wa * canvasA[:, both] + wb * canvasB[:, both]
    # astype(np.uint8): the C cast (truncate, wrap modulo 256 for out-of-range values)
    return result.astype(np.int32).astype(np.uint8)


def rectangle_crop(img, black_threshold, extra_margin):
    """image_stitching_sift.py:208-247."""
    h, w = img.shape[:2]
    gray = bgr2gray(img)
    ys, xs = np.where(gray > black_threshold)
    if ys.size == 0:
        return img
    y_min, y_max, x_min, x_max = ys.min(), ys.max(), xs.min(), xs.max()
    y_min = max(0, y_min + extra_margin)
    y_max = min(h - 1, y_max - extra_margin)
    if y_min > y_max or x_min > x_max:
        return img
    return img[y_min:y_max + 1, x_min:x_max + 1]


def drift_corrected_shifts(shift_list, n_images):
    """image_stitching_sift.py:336-365: every dy loses the average drift final_dy / (N - 1)."""
    acc = [(0, 0)]
    for i in range(len(shift_list)):
        acc.append((acc[i][0] + shift_list[i][0], acc[i][1] + shift_list[i][1]))
    average_drift = acc[-1][1] / (n_images - 1) if n_images > 1 else 0
    return [(dx, dy - average_drift) for dx, dy in shift_list]


def stitch(cyl_imgs, shift_list, matched_pairs):
    """Second loop of run_panorama (image_stitching_sift.py:367-381) + the default crop (:383-384)."""
    new_shifts = drift_corrected_shifts(shift_list, len(cyl_imgs))
    mosaic = cyl_imgs[0].copy()
    for i in range(1, len(cyl_imgs)):
        img = cyl_imgs[i]
        diff_y = mosaic.shape[0] - img.shape[0]
        if diff_y != 0:
            img = pad_image(img, 0, diff_y)
        mosaic = blend_two_images(new_shifts[i - 1], matched_pairs[i - 1], mosaic, img)
    return mosaic
