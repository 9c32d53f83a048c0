"""ctypes binding of libb200sift.so (include/b200sift.h).

This is the only place the package touches the native library.  There is no
CPU fallback: if the library is missing, cannot be loaded, or no sm_100 GPU
is present, the calls raise.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200sift.so')

# numpy mirror of b200sift_keypoint (== the cv2.KeyPoint fields the reference uses)
KP_DTYPE = np.dtype([('x', np.float32), ('y', np.float32), ('size', np.float32), ('angle', np.float32),
                     ('response', np.float32), ('octave', np.int32)])


class Params(C.Structure):
    """b200sift_params: the reference's keyword defaults (sift_impl.py:15,117,170,247,361-362)."""
    _fields_ = [('sigma', C.c_double), ('num_intervals', C.c_int32), ('assumed_blur', C.c_double),
                ('image_border_width', C.c_int32), ('contrast_threshold', C.c_double),
                ('eigen_ratio', C.c_double), ('max_iter', C.c_int32), ('radius_factor', C.c_double),
                ('ori_bins', C.c_int32), ('peak_ratio', C.c_double), ('scale_factor', C.c_double),
                ('window_width', C.c_int32), ('desc_bins', C.c_int32), ('scale_multiplier', C.c_double),
                ('descriptor_max_value', C.c_double)]


class B200SiftError(RuntimeError):
    pass


_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_ip = C.POINTER(C.c_int32)
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol include/b200sift.h declares
PROTOTYPES = {
    'b200sift_default_params': (None, [C.POINTER(Params)]),
    'b200sift_last_error': (C.c_char_p, []),
    'b200sift_version': (C.c_char_p, []),
    'b200sift_create': (_i, [_i, _pp]),
    'b200sift_destroy': (None, [_vp]),
    'b200sift_set_stream': (_i, [_vp, _vp]),
    'b200sift_get_stream': (_i, [_vp, _pp]),
    'b200sift_set_sync_mode': (_i, [_vp, _i]),
    'b200sift_last_kernel_ms': (_i, [_vp, C.POINTER(C.c_float)]),
    'b200sift_last_describe_ms': (_i, [_vp, C.POINTER(C.c_float), _ip]),
    'b200sift_launch_count': (_i, [_vp, C.POINTER(C.c_longlong)]),
    'b200sift_sync': (_i, [_vp]),
    'b200sift_detect_describe': (_i, [_vp, C.POINTER(Params), _i, _pp, _i, _i, _i, _i, _sz, _i, _ip]),
    'b200sift_get_keypoints': (_i, [_vp, _i, _vp, _vp, _vp]),
    'b200sift_get_all_keypoints': (_i, [_vp, _vp, _vp, C.c_int64]),
    'b200sift_get_stats': (_i, [_vp, _i, _ip, _ip, _ip]),
    'b200sift_device_results': (_i, [_vp, _i, _pp, _pp, _ip]),
    'b200sift_match': (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'b200sift_match_images': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _ip]),
    'b200sift_match_pairs': (_i, [_vp, _i, _ip, _i, _d, C.POINTER(C.c_double), _ip, _ip, _vp]),
    'b200sift_get_pair_matches': (_i, [_vp, _i, _vp, _vp, _vp]),
    'b200sift_append_results': (_i, [_vp, _vp, _vp, _i, _i, _ip]),
    'b200sift_pack_exchange': (_i, [_vp, _i, _ip, _i, _vp, _i]),
    'b200sift_unpack_exchange': (_i, [_vp, _vp, _i, _i, _i, _ip, _ip]),
    'b200sift_append_exchange': (_i, [_vp, _vp, _i, _ip]),
    'b200sift_match_pairs_device': (_i, [_vp, _i, _ip, _i, _d, _vp, _sz]),
    'b200sift_blend_two_images': (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _d, _d, C.POINTER(C.c_double), _i, _vp, _sz,
                                        _ip, _ip]),
    'b200sift_crop_bbox': (_i, [_vp, _vp, _i, _i, _i, _ip]),
    'b200sift_ransac': (_i, [_vp, _vp, _i, _d, C.POINTER(C.c_double), _ip]),
    'b200sift_gaussian_blur': (_i, [_vp, _vp, _i, _i, _d, _vp, _i]),
    'b200sift_base_image': (_i, [_vp, _vp, _i, _i, _d, _d, _vp]),
    'b200sift_gaussian_pyramid': (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_double), _i, _pp]),
    'b200sift_dog_pyramid': (_i, [_vp, _pp, _i, _i, _i, _i, _pp]),
    'b200sift_find_extrema': (_i, [_vp, C.POINTER(Params), _pp, _pp, _i, _i, _i, _i, _vp, _i, _ip]),
    'b200sift_localize': (_i, [_vp, C.POINTER(Params), _pp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    'b200sift_orientations': (_i, [_vp, C.POINTER(Params), _vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    'b200sift_ratio_match': (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _ip]),
    'b200sift_match_grid': (_i, [_vp, _ip, _ip]),
    'b200sift_extrema_candidates': (_i, [_vp, C.POINTER(Params), _pp, _i, _i, _i, _i, _vp, _i, _ip]),
    'b200sift_remove_duplicates': (_i, [_vp, _vp, _i, _ip]),
    'b200sift_descriptors': (_i, [_vp, C.POINTER(Params), _vp, _i, _pp, _i, _i, _i, _i, _vp]),
    'b200sift_cylindrical_projection': (_i, [_vp, _vp, _i, _i, _i, _d, _vp]),
    'b200sift_bench_match': (_i, [_vp, _vp, _i, _vp, _i, _i, _i, C.POINTER(C.c_float)]),
    'b200sift_bench_blur': (_i, [_vp, _i, _i, _i, _d, _i, _i, C.POINTER(C.c_float)]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libb200sift.so and bind every prototype.  Raises if the library is absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise B200SiftError(
                    f'{LIB_PATH} is missing: build it with `python -m vfx_image_stitching_b200.build` '
                    '(nvcc, sm_100a). There is no CPU fallback.')
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise B200SiftError(f'b200sift error {rc}: {load().b200sift_last_error().decode(errors="replace")}')


def default_params(**overrides):
    p = Params()
    load().b200sift_default_params(C.byref(p))
    for k, v in overrides.items():
        if v is not None:
            setattr(p, k, v)
    return p


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def ptr_array(arrays):
    """void*[] over a list of numpy arrays (the arrays must stay referenced by the caller)."""
    arr = (C.c_void_p * max(1, len(arrays)))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr


class Context:
    """One b200sift context = one GPU.  Not thread-safe."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.lib = load()
        check(self.lib.b200sift_create(int(device), C.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        if not self._h:
            raise B200SiftError('context already destroyed')
        return self._h

    def close(self):
        if self._h:
            self.lib.b200sift_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        check(self.lib.b200sift_set_stream(self.handle, C.c_void_p(cuda_stream_ptr or 0)))

    def set_sync_mode(self, mode):
        """0 spin, 1 poll + yield, 2 sleep while a synchronous call waits for the device."""
        check(self.lib.b200sift_set_sync_mode(self.handle, int(mode)))

    def stream_handle(self):
        """cudaStream_t (int) the context launches on."""
        st = C.c_void_p()
        check(self.lib.b200sift_get_stream(self.handle, C.byref(st)))
        return st.value or 0

    def last_kernel_ms(self):
        ms = C.c_float()
        check(self.lib.b200sift_last_kernel_ms(self.handle, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_longlong()
        check(self.lib.b200sift_launch_count(self.handle, C.byref(n)))
        return n.value

    def sync(self):
        check(self.lib.b200sift_sync(self.handle))


_default = {}


def default_context(device=None):
    """Process-wide context for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    ctx = _default.get(device)
    if ctx is None or not ctx._h:
        ctx = Context(device)
        _default[device] = ctx
    return ctx
