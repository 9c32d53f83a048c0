"""Row f4 on the GPU: blend_two_images / rectangle_crop / the second loop of run_panorama through the
C ABI against the golden vectors of the unmodified reference (bit-exact: byte work)."""
import numpy as np
import pytest

from test_oracle_stitch import GOLD, blend_case, set_inputs, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


@pytest.fixture(scope='module')
def iss():
    from vfx_image_stitching_b200 import image_stitching_sift
    return image_stitching_sift


def test_blend_cases_bit_exact(gold, iss):
    from oracle import sift_oracle as so
    for i in range(int(gold['n_blend'])):
        shift, rm, a, b, want, _ = blend_case(gold, i)
        got = iss.blend_two_images(shift, rm, a, b)
        assert got.shape == want.shape and np.array_equal(got, want), i
        assert np.array_equal(got, so.blend_two_images(shift, rm, a, b)), i


def test_blend_random_against_oracle(iss):
    """shapes, shifts and black regions the golden file does not hold; both alpha dtypes"""
    from oracle import sift_oracle as so
    rng = np.random.default_rng(3)
    for k in range(12):
        ha, wa = int(rng.integers(8, 70)), int(rng.integers(12, 90))
        hb, wb = ha + int(rng.integers(0, 3)), int(rng.integers(12, 90))
        a = rng.integers(0, 256, (ha, wa, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (hb, wb, 3), dtype=np.uint8)
        a[:, :int(rng.integers(0, 4))] = 0
        b[:, wb - int(rng.integers(1, 5)):] = 0
        dx, dy = float(rng.uniform(-wa, wa)), float(rng.uniform(-6, 6))
        xa, ya = float(rng.uniform(0, wa)), float(rng.uniform(0, ha))
        cast = np.float64 if k % 2 else float
        rm = ((cast(xa), cast(ya)), (cast(xa - dx), cast(ya - dy)))
        assert np.array_equal(iss.blend_two_images((dx, dy), rm, a, b), so.blend_two_images((dx, dy), rm, a, b)), k


def test_crop_cases(gold, iss):
    for i in range(int(gold['n_crop'])):
        got = iss.rectangle_crop(gold[f'crop{i}_img'], int(gold[f'crop{i}_thr']), int(gold[f'crop{i}_margin']))
        assert np.array_equal(got, gold[f'crop{i}_out']), i


@pytest.mark.parametrize('name', ['out', 'parrington', 'grail'])
def test_second_loop_matches_reference(name, gold, iss):
    """the reference's shifts and matched pairs in, its mosaic and crop out (SHA-256 of the bytes)"""
    cyl, shifts, pairs = set_inputs(name)
    new_shifts = iss.drift_corrected_shifts(shifts, len(cyl))
    mosaic = cyl[0].copy()
    for i in range(1, len(cyl)):
        nxt = cyl[i]
        if mosaic.shape[0] != nxt.shape[0]:
            nxt = iss.pad_image(nxt, 0, mosaic.shape[0] - nxt.shape[0])
        mosaic = iss.blend_two_images(new_shifts[i - 1], pairs[i - 1], mosaic, nxt)
    assert list(mosaic.shape) == list(gold[f'{name}_mosaic_shape'])
    assert sha(mosaic) == str(gold[f'{name}_mosaic_sha'])
    crop = iss.rectangle_crop(mosaic, 0, 15)
    assert list(crop.shape) == list(gold[f'{name}_crop_shape'])
    assert sha(crop) == str(gold[f'{name}_crop_sha'])


def test_stitch_panorama_end_to_end(gold, iss):
    """projected images in, panorama out, with the GPU's own SIFT shifts: the shifts agree with the
    reference's to < 1e-3 px, so the pads round the same way and the mosaic has the same shape; the
    cross-fade weights move by ~1e-6, i.e. a grey level may flip in a few pixels."""
    cyl, shifts, _ = set_inputs('out')
    result, mosaic, got_shifts, _ = iss.stitch_panorama(cyl)
    assert max(abs(a - b) for s, t in zip(got_shifts, shifts) for a, b in zip(s, t)) < 1e-3
    want = gold['out_mosaic']
    assert mosaic.shape == want.shape
    diff = np.abs(mosaic.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and np.mean(diff > 0) < 1e-3
    assert list(result.shape) == list(gold['out_crop_shape'])
