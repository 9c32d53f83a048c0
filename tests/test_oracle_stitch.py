"""Row f4 of the scope table: the oracle's restatement of read_pano_data / pad_image /
blend_two_images / rectangle_crop / the drift-corrected second loop of run_panorama against
golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden_stitch.py)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, 'tests', 'golden', 'stitch.npz')


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


@pytest.fixture(scope='module')
def so():
    from oracle import sift_oracle
    return sift_oracle


def blend_case(g, i):
    pair = g[f'blend{i}_pair']
    strong = bool(g[f'blend{i}_strong'])
    cast = np.float64 if strong else float
    rm = ((cast(pair[0]), cast(pair[1])), (cast(pair[2]), cast(pair[3])))
    shift = tuple(float(v) for v in g[f'blend{i}_shift'])
    return shift, rm, g[f'blend{i}_a'], g[f'blend{i}_b'], g[f'blend{i}_out'], strong


def set_inputs(name):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    if 'bgr_0' in g.files and all(f'bgr_{i}' in g.files for i in range(len(g['gray']))):
        cyl = [g[f'bgr_{i}'].copy() for i in range(len(g['gray']))]
    else:
        cyl = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g['gray']]
    shifts = [tuple(float(v) for v in s) for s in g['shifts']]
    pairs = [((float(b[0]), float(b[1])), (float(b[2]), float(b[3]))) for b in g['best_pairs']]
    return cyl, shifts, pairs


def test_blend_cases_bit_exact(gold, so):
    for i in range(int(gold['n_blend'])):
        shift, rm, a, b, want, _ = blend_case(gold, i)
        got = so.blend_two_images(shift, rm, a, b)
        assert got.shape == want.shape and np.array_equal(got, want), i


def test_crop_cases(gold, so):
    for i in range(int(gold['n_crop'])):
        got = so.rectangle_crop(gold[f'crop{i}_img'], int(gold[f'crop{i}_thr']), int(gold[f'crop{i}_margin']))
        assert np.array_equal(got, gold[f'crop{i}_out']), i


def test_read_pano_data(tmp_path, gold, so):
    # the reference's own pano.txt layout: path line, blank/matrix lines, focal line
    lines = []
    for p, f in zip(gold['out_paths'], gold['out_focals']):
        lines += [str(p), '571 428', '', '1 0 0', '0 1 0', '0 0 1', '', repr(float(f)), '']
    lines += ['no_image_here', '12.5']        # a focal without a pending image is ignored
    fn = tmp_path / 'pano.txt'
    fn.write_text('\n'.join(lines), encoding='utf-8')
    paths, focals = so.read_pano_data(str(fn))
    assert paths == [str(p) for p in gold['out_paths']]
    assert focals == [float(f) for f in gold['out_focals']]


@pytest.mark.parametrize('name', ['out', 'parrington', 'grail'])
def test_second_loop_matches_reference(name, gold, so):
    cyl, shifts, pairs = set_inputs(name)
    mosaic = so.stitch(cyl, shifts, pairs)
    assert list(mosaic.shape) == list(gold[f'{name}_mosaic_shape'])
    assert sha(mosaic) == str(gold[f'{name}_mosaic_sha'])
    if name == 'out':
        assert np.array_equal(mosaic, gold['out_mosaic'])
    crop = so.rectangle_crop(mosaic, 0, 15)
    assert list(crop.shape) == list(gold[f'{name}_crop_shape'])
    assert sha(crop) == str(gold[f'{name}_crop_sha'])
