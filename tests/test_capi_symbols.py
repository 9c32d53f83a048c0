"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/b200sift.h declares; the ctypes layer binds all of them; the product path fails
loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, _has_gpu


def _declared():
    src = open(os.path.join(ROOT, 'include', 'b200sift.h')).read()
    return sorted(set(re.findall(r'B200SIFT_API[^;(]*?\b(b200sift_\w+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    from vfx_image_stitching_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/b200sift.h but not exported'


def test_integration_doc_names_every_entry_point():
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing


def test_f4_signatures_and_host_logic(tmp_path):
    """Row f4 mirrors: names / argument order of image_stitching_sift.py:12,139,156,208 and the pure host
    logic (pano.txt parser, padding, drift correction) against the oracle's restatement -- no GPU involved."""
    import inspect
    from oracle import sift_oracle as so
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    names = lambda f: [p.name for p in inspect.signature(f).parameters.values()]  # noqa: E731
    assert names(iss.read_pano_data) == ['pano_file_path']
    assert names(iss.pad_image) == ['img_bgr', 'move_x', 'move_y']
    assert names(iss.blend_two_images)[:4] == ['shift_vec', 'ref_match', 'imgA', 'imgB']
    assert names(iss.rectangle_crop)[:3] == ['img', 'black_threshold', 'extra_margin']
    assert names(iss.cylindrical_projection)[:2] == ['img_bgr', 'focal_len']
    fn = tmp_path / 'pano.txt'
    fn.write_text('C:\\x\\a01.JPG\n384 512\n\n1 0 0\n0 1 0\n0 0 1\n\n704.5\nb02.png\nnot a number\n 12 \n703\n9.5\n',
                  encoding='utf-8')
    assert iss.read_pano_data(str(fn)) == so.read_pano_data(str(fn)) == (['C:\\x\\a01.JPG', 'b02.png'], [704.5, 12.0])
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    for mx, my in ((2.5, -1.5), (-3.4, 0.5), (0, 0), (1.5, 2.5)):
        assert np.array_equal(iss.pad_image(img, mx, my), so.pad_image(img, mx, my))
    shifts = [(-240.5, -4.25), (-251.0, 3.5), (-239.75, -6.0)]
    assert iss.drift_corrected_shifts(shifts, 4) == so.drift_corrected_shifts(shifts, 4)


def test_ctypes_prototypes_cover_header():
    from vfx_image_stitching_b200 import _capi
    assert sorted(_capi.PROTOTYPES) == _declared()
    lib = _capi.load()
    assert lib.b200sift_version().decode().startswith('b200sift')
    p = _capi.default_params()
    assert (p.sigma, p.num_intervals, p.assumed_blur, p.image_border_width) == (1.6, 3, 0.5, 5)
    assert (p.contrast_threshold, p.eigen_ratio, p.max_iter) == (0.04, 10, 5)
    assert (p.radius_factor, p.ori_bins, p.peak_ratio, p.scale_factor) == (3, 36, 0.8, 1.5)
    assert (p.window_width, p.desc_bins, p.scale_multiplier, p.descriptor_max_value) == (4, 8, 3, 0.2)
    assert _capi.KP_DTYPE.itemsize == 24


def test_reference_signatures_are_preserved():
    """Names, positional order and defaults of sift_impl.py / image_stitching_sift.py."""
    import inspect
    from vfx_image_stitching_b200 import sift_impl, image_stitching_sift

    def sig(f):
        return [(p.name, p.default if p.default is not inspect._empty else None)
                for p in inspect.signature(f).parameters.values()]
    assert sig(sift_impl.compute_keypoints_and_descriptors) == [
        ('image', None), ('sigma', 1.6), ('num_intervals', 3), ('assumed_blur', 0.5), ('image_border_width', 5)]
    assert sig(sift_impl.generate_base_image) == [('image', None), ('sigma', None), ('assumed_blur', None)]
    assert sig(sift_impl.generate_gaussian_images) == [('image', None), ('num_octaves', None), ('gaussian_kernels', None)]
    assert sig(sift_impl.find_scale_space_extrema) == [
        ('gaussian_images', None), ('dog_images', None), ('num_intervals', None), ('sigma', None), ('border', None),
        ('contrast_threshold', 0.04)]
    assert sig(sift_impl.generate_descriptors) == [
        ('keypoints', None), ('gaussian_images', None), ('window_width', 4), ('num_bins', 8),
        ('scale_multiplier', 3), ('descriptor_max_value', 0.2)]
    assert sig(sift_impl.localize_extremum_via_quadratic_fit) == [
        ('x', None), ('y', None), ('layer', None), ('octave', None), ('num_intervals', None), ('dog_octave', None),
        ('sigma', None), ('contrast_threshold', None), ('border', None), ('eigen_ratio', 10), ('max_iter', 5)]
    assert sig(sift_impl.compute_keypoints_with_orientations) == [
        ('keypoint', None), ('octave', None), ('gauss_img', None), ('radius_factor', 3), ('num_bins', 36),
        ('peak_ratio', 0.8), ('scale_factor', 1.5)]
    assert sig(sift_impl.is_pixel_an_extremum) == [
        ('prev_patch', None), ('curr_patch', None), ('next_patch', None), ('threshold', None)]
    assert sig(image_stitching_sift.compute_shift_sift)[:4] == [
        ('imgA', None), ('imgB', None), ('ransac_thr', 3), ('desc_thresh', 25000)]
    assert sig(image_stitching_sift.ransac)[:2] == [('matches', None), ('dist_sq_thresh', 3)]
    assert sift_impl.float_tolerance == 1e-7
    for name in ('compute_number_of_octaves', 'generate_gaussian_kernels', 'generate_DoG_images',
                 'is_pixel_an_extremum', 'compute_gradient_at_center_pixel', 'compute_hessian_at_center_pixel',
                 'compare_keypoints', 'remove_duplicate_keypoints', 'convert_keypoints_to_input_image_size',
                 'unpack_octave'):
        assert callable(getattr(sift_impl, name))


def test_every_reference_function_is_mirrored():
    """All 17 top-level functions of the reference's sift_impl.py exist in the drop-in with the
    same parameter names in the same order; checked against the list recorded from the reference source."""
    import inspect
    from vfx_image_stitching_b200 import sift_impl
    ref = {  # name -> positional parameter names, /root/reference/sift_impl.py (def lines 15 ... 361)
        'compute_keypoints_and_descriptors': ['image', 'sigma', 'num_intervals', 'assumed_blur', 'image_border_width'],
        'generate_base_image': ['image', 'sigma', 'assumed_blur'],
        'compute_number_of_octaves': ['image_shape'],
        'generate_gaussian_kernels': ['sigma', 'num_intervals'],
        'generate_gaussian_images': ['image', 'num_octaves', 'gaussian_kernels'],
        'generate_DoG_images': ['gaussian_images'],
        'find_scale_space_extrema': ['gaussian_images', 'dog_images', 'num_intervals', 'sigma', 'border',
                                     'contrast_threshold'],
        'is_pixel_an_extremum': ['prev_patch', 'curr_patch', 'next_patch', 'threshold'],
        'localize_extremum_via_quadratic_fit': ['x', 'y', 'layer', 'octave', 'num_intervals', 'dog_octave', 'sigma',
                                                'contrast_threshold', 'border', 'eigen_ratio', 'max_iter'],
        'compute_gradient_at_center_pixel': ['cube'],
        'compute_hessian_at_center_pixel': ['cube'],
        'compute_keypoints_with_orientations': ['keypoint', 'octave', 'gauss_img', 'radius_factor', 'num_bins',
                                                'peak_ratio', 'scale_factor'],
        'compare_keypoints': ['kp1', 'kp2'],
        'remove_duplicate_keypoints': ['keypoints'],
        'convert_keypoints_to_input_image_size': ['keypoints'],
        'unpack_octave': ['keypoint'],
        'generate_descriptors': ['keypoints', 'gaussian_images', 'window_width', 'num_bins', 'scale_multiplier',
                                 'descriptor_max_value'],
    }
    assert len(ref) == 17
    for name, params in ref.items():
        assert [p.name for p in inspect.signature(getattr(sift_impl, name)).parameters.values()] == params, name


def test_desc_thresh_is_the_reference_comparison():
    """`dist < desc_thresh` on integer distances (image_stitching_sift.py:74) for non-integer thresholds."""
    from vfx_image_stitching_b200.image_stitching_sift import _int_thresh
    assert _int_thresh(25000) == 25000 and _int_thresh(25000.5) == 25001 and _int_thresh(24999.0001) == 25000
    assert _int_thresh(-3.5) == -3 and _int_thresh(1e12) == 2 ** 31 - 1


def test_host_side_parameter_arithmetic_matches_oracle(oracle):
    from vfx_image_stitching_b200 import sift_impl
    assert np.array_equal(sift_impl.generate_gaussian_kernels(1.6, 3), oracle.generate_gaussian_kernels(1.6, 3))
    for shape in ((1024, 768), (1142, 856), (8192, 6144), (10, 14), (4, 4)):
        assert sift_impl.compute_number_of_octaves(shape) == oracle.compute_number_of_octaves(shape)
    k = sift_impl.array_to_keypoints(np.array([(3.5, 4.25, 2.0, 90.0, 0.1, 0x2FF)], sift_impl.KP_DTYPE))[0]
    assert sift_impl.unpack_octave(k)[:2] == (-1, 2) and sift_impl.unpack_octave(k)[2] == 2.0
    k.octave = 0x10301
    sift_impl.convert_keypoints_to_input_image_size([k])
    assert k.pt == (1.75, 2.125) and k.size == 1.0 and k.octave == 0x10300


@pytest.mark.skipif(_has_gpu(), reason='checks the behaviour WITHOUT a GPU')
def test_no_cpu_fallback_without_gpu():
    from vfx_image_stitching_b200 import sift_impl
    from vfx_image_stitching_b200._capi import B200SiftError
    with pytest.raises(B200SiftError):
        sift_impl.compute_keypoints_and_descriptors(np.zeros((32, 32), np.uint8))


@pytest.mark.skipif(_has_gpu(), reason='checks the behaviour WITHOUT a GPU')
def test_pipeline_needs_a_gpu():
    from vfx_image_stitching_b200._capi import B200SiftError
    from vfx_image_stitching_b200.pipeline import PanoramaPipeline
    with pytest.raises(B200SiftError):
        PanoramaPipeline(depth=2)
