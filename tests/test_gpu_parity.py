"""Parity of the CUDA path (through the C ABI) against the oracle and the reference goldens.

Run on the GPU box: python -m pytest tests -m gpu.  Tolerances:
  * integer / index work (grey, upsample of integer images, extrema candidates, ordering,
    de-duplication, matcher, vote, projection): bit-exact;
  * blur: <= 2e-4 abs on the 0..255 range against the oracle AND against cv2.GaussianBlur (the
    reference's own IPP-on vs IPP-off blur differs by 7.6e-5);
  * keypoints given the same pyramid: x, y, size, response, octave bit-exact (float64 solve,
    no FMA), angle <= 1e-3 deg for >= 99.5 %;
  * descriptors given the same keypoints + pyramid: |diff| <= 1 quantisation step, >= 97 % of rows
    identical (float32 accumulation order differs from np.add.at's sequential order);
  * end to end against the reference: north_star (>= 99 % keypoints within 0.5 px / 0.05 octave,
    descriptor RMS relative L2, shift within 0.5 px).
"""
import os

import numpy as np
import pytest

from conftest import ROOT, golden_kps, match_keypoint_sets, natural_image

pytestmark = pytest.mark.gpu

REPORT = os.path.join(ROOT, 'gpurun_out', 'parity_report.txt')


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, 'a') as f:
        f.write(line + '\n')
    print(line)


@pytest.fixture(scope='module')
def si():
    from vfx_image_stitching_b200 import sift_impl
    return sift_impl


@pytest.fixture(scope='module')
def iss():
    from vfx_image_stitching_b200 import image_stitching_sift
    return image_stitching_sift


SIGMAS = [1.2489996, 1.2262735, 1.5450078, 1.9465878, 2.452547, 3.0900156]


# ----------------------------------------------------------------------------- dense stage
@pytest.mark.parametrize('shape', [(1024, 768), (571, 428), (1142, 856), (96, 128), (64, 200), (33, 97), (12, 9),
                                   (4, 3), (1, 7), (300, 1030)])
def test_blur_matches_oracle_and_cv2(si, oracle, shape):
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = (rng.random(shape) * 255).astype(np.float32)
    for s in SIGMAS + [0.7, 4.3]:
        got = si.gaussian_blur(img, s)
        ref = oracle.gaussian_blur(img, s, 'c')
        err = float(np.abs(got - ref).max())
        err_cv = float(np.abs(got - cv2.GaussianBlur(img, (0, 0), sigmaX=s, sigmaY=s)).max())
        report(f'blur {shape} sigma={s:.4f} max|gpu-oracle|={err:.2e} max|gpu-cv2|={err_cv:.2e}')
        assert err < 2e-4 and err_cv < 2e-4


def test_blur_ring_edges_and_determinism(si, oracle):
    """The packed ring kernel's border patch, partial last strip / last batch and its mbarrier
    hand-off: widths and heights around the 256-column strip, the 8-row batch and the 4-float chunk;
    every result must repeat bit for bit (a race between the producer and consumer warps would not)."""
    rng = np.random.default_rng(99)
    for h, w in [(32, 96), (33, 97), (39, 255), (40, 256), (41, 257), (64, 259), (47, 511), (130, 513),
                 (257, 770), (96, 1026)]:
        img = (rng.random((h, w)) * 255).astype(np.float32)
        for s in (1.2262735, 2.452547, 3.0900156):
            got = si.gaussian_blur(img, s)
            ref = oracle.gaussian_blur(img, s, 'c')
            assert np.abs(got - ref).max() < 2e-4, (h, w, s)
            for _ in range(3):
                assert np.array_equal(si.gaussian_blur(img, s), got), (h, w, s)


def test_base_image_bit_exact_upsample(si, oracle):
    g = natural_image(120, 90, 1)
    got = si.generate_base_image(g.astype(np.float32), 1.6, 0.5)
    ref = oracle.generate_base_image(g.astype(np.float32), 1.6, 0.5)
    assert got.shape == (240, 180)
    assert np.abs(got - ref).max() < 2e-4
    # sigma=None -> 1.6 (sift_impl.py:50-51)
    assert np.array_equal(si.generate_base_image(g.astype(np.float32), None, 0.5), got)


@pytest.mark.parametrize('shape', [(240, 180), (1024, 768), (70, 50)])
def test_pyramid_and_dog(si, oracle, shape):
    rng = np.random.default_rng(5)
    base = (rng.random(shape) * 255).astype(np.float32)
    n_oct = si.compute_number_of_octaves(base.shape)
    sig = si.generate_gaussian_kernels(1.6, 3)
    got = si.generate_gaussian_images(base, n_oct, sig)
    ref = oracle.generate_gaussian_images(base, n_oct, sig)
    assert got.shape == (n_oct, 6) and got.dtype == object
    worst = 0.0
    for o in range(n_oct):
        for l in range(6):
            assert got[o, l].shape == ref[o][l].shape and got[o, l].dtype == np.float32
            worst = max(worst, float(np.abs(got[o, l] - ref[o][l]).max()))
    report(f'pyramid {shape} octaves={n_oct} max|gpu-oracle|={worst:.2e}')
    assert worst < 5e-4
    assert np.array_equal(got[0, 0], base)
    assert np.array_equal(got[1, 0], got[0, 3][::2, ::2][:shape[0] // 2, :shape[1] // 2])
    dog = si.generate_DoG_images(got)
    assert dog.shape == (n_oct, 5)
    for o in range(n_oct):
        for l in range(5):
            assert np.array_equal(dog[o, l], got[o, l + 1] - got[o, l])       # exact float32 subtraction


# ----------------------------------------------------------------------------- sparse stage on a given pyramid
@pytest.fixture(scope='module')
def ref_pyramid(oracle, golden):
    """Pyramid of out[1] built with the oracle's blur: the common input of the stage tests."""
    gray = golden('out')['gray'][1].astype(np.float32)
    base = oracle.generate_base_image(gray, 1.6, 0.5)
    g = oracle.generate_gaussian_images(base, oracle.compute_number_of_octaves(base.shape),
                                        oracle.generate_gaussian_kernels(1.6, 3))
    return g


def test_extrema_candidates_bit_exact(si, oracle, ref_pyramid):
    _, st = oracle.find_scale_space_extrema(ref_pyramid, None, 3, 1.6, 5, return_stats=True)
    got = si.extrema_candidates(ref_pyramid)
    report(f'candidates gpu={len(got)} oracle={len(st["candidates"])}')
    assert np.array_equal(got, st['candidates'])


def test_find_extrema_matches_oracle(si, oracle, ref_pyramid):
    ref = oracle.find_scale_space_extrema(ref_pyramid, None, 3, 1.6, 5)
    got = si.find_scale_space_extrema_array(ref_pyramid)
    report(f'find_extrema gpu={len(got)} oracle={len(ref)}')
    assert len(got) == len(ref)
    for f in ('x', 'y', 'size', 'response', 'octave'):
        same = float(np.mean(got[f] == ref[f]))
        report(f'   field {f}: identical {same:.4f}')
        assert same >= 0.998, f     # float64 direct solve vs the oracle's pseudo-inverse: last-bit only
    dang = np.abs(got['angle'] - ref['angle'])
    dang = np.minimum(dang, 360 - dang)
    report(f'   angle identical {np.mean(dang == 0):.4f} max {dang.max():.3e}')
    assert np.mean(dang < 1e-3) >= 0.995
    kps = si.find_scale_space_extrema(ref_pyramid, None, num_intervals=3, sigma=1.6, border=5)
    assert len(kps) == len(ref) and abs(kps[0].pt[0] - ref['x'][0]) == 0


def test_remove_duplicates_bit_exact(si, oracle, ref_pyramid):
    raw = oracle.find_scale_space_extrema(ref_pyramid, None, 3, 1.6, 5)
    rng = np.random.default_rng(0)
    dup = np.concatenate([raw, raw[rng.integers(0, len(raw), 200)]])   # force repeats
    dup = dup[rng.permutation(len(dup))]
    ref = oracle.remove_duplicate_keypoints(dup)
    got = si.keypoints_to_array(si.remove_duplicate_keypoints(si.array_to_keypoints(dup)))
    assert len(got) == len(ref)
    for f in ref.dtype.names:
        assert np.array_equal(got[f], ref[f]), f
    one = si.array_to_keypoints(raw[:1])
    assert si.remove_duplicate_keypoints(one) is one              # < 2 returned as is (:318-319)


def test_descriptors_match_oracle(si, oracle, ref_pyramid):
    raw = oracle.find_scale_space_extrema(ref_pyramid, None, 3, 1.6, 5)
    kps = oracle.convert_keypoints_to_input_image_size(oracle.remove_duplicate_keypoints(raw))
    ref = oracle.generate_descriptors(kps, ref_pyramid)
    got = si.generate_descriptors(si.array_to_keypoints(kps), ref_pyramid)
    assert got.shape == ref.shape and got.dtype == np.float32
    d = np.abs(got - ref)
    same = float(np.mean(d.sum(1) == 0))
    report(f'descriptors n={len(kps)} identical rows {same:.4f} max|diff| {d.max():.0f} '
           f'mean flips/row {np.mean((d > 0).sum(1)):.4f}')
    assert d.max() <= 1
    assert same >= 0.97
    assert si.generate_descriptors([], ref_pyramid).shape == (0,)


def test_pairs_batched_equals_single_calls(iss, si, golden):
    g = golden('grail')
    imgs = [g['gray'][i] for i in range(3)]
    res = si.detect_and_describe_batch(imgs)
    shifts, nm, best, bp = iss.match_pairs([(0, 1), (1, 2), (2, 0), (1, 1)])
    for p, (a, b) in enumerate([(0, 1), (1, 2), (2, 0), (1, 1)]):
        ka, kb = si.array_to_keypoints(res[a][0]), si.array_to_keypoints(res[b][0])
        ia, ib, matches = iss.match_keypoints(ka, res[a][1], kb, res[b][1])
        assert nm[p] == len(ia)
        move, pair = iss.ransac(matches, 3)
        assert shifts[p] == tuple(move) and bp[p] == pair
    assert shifts[3] == (0.0, 0.0) and nm[3] == len(res[1][0])      # an image against itself


# ----------------------------------------------------------------------------- end to end vs the reference
@pytest.mark.parametrize('name,idx', [('out', 0), ('out', 1), ('parrington', 0), ('parrington', 1), ('grail', 2)])
def test_end_to_end_against_reference_golden(si, golden, name, idx):
    g = golden(name)
    (kps, desc), = si.detect_and_describe_batch([g['gray'][idx]])
    ref = golden_kps(g, idx)
    frac, m = match_keypoint_sets(ref, kps)
    a = g[f'desc_{idx}'][m >= 0].astype(np.float64)
    b = desc[m[m >= 0]].astype(np.float64)
    rel = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(a, axis=1), 1)
    dang = np.abs(ref['angle'][m >= 0] - kps['angle'][m[m >= 0]])
    dang = np.minimum(dang, 360 - dang)
    report(f'e2e {name}[{idx}] ref={len(ref)} gpu={len(kps)} matched={frac:.4f} desc rms rel-L2='
           f'{np.sqrt(np.mean(rel ** 2)):.2e} median={np.median(rel):.1e} rows identical='
           f'{np.mean(rel == 0):.3f} angle<1deg={np.mean(dang < 1):.4f} stats={si.stage_stats(0)}')
    assert frac >= 0.99
    assert abs(len(kps) - len(ref)) <= max(3, len(ref) // 100)
    assert np.median(rel) < 1e-3 and np.sqrt(np.mean(rel ** 2)) < 1e-3     # north_star: 1e-3 relative L2
    # ordering contract of remove_duplicate_keypoints
    key = np.stack([kps['x'], kps['y']], 1)
    assert np.all((key[1:, 0] > key[:-1, 0]) | ((key[1:, 0] == key[:-1, 0]) & (key[1:, 1] >= key[:-1, 1])))


def test_bgr_and_float_inputs_agree(si, golden):
    g = golden('out')
    bgr = g['bgr_0']
    (k1, d1), = si.detect_and_describe_batch([bgr])
    (k2, d2), = si.detect_and_describe_batch([g['gray'][0]])
    (k3, d3), = si.detect_and_describe_batch([g['gray'][0].astype(np.float64)])
    for k, d in ((k2, d2), (k3, d3)):
        assert np.array_equal(k1, k) and np.array_equal(d1, d)
    kps, desc = si.compute_keypoints_and_descriptors(bgr)
    assert len(kps) == len(k1) and desc.dtype == np.float32 and desc.shape == (len(k1), 128)
    assert kps[0].pt == (float(k1['x'][0]), float(k1['y'][0])) and kps[0].octave == int(k1['octave'][0])


def test_batch_equals_single(si, golden):
    g = golden('parrington')
    imgs = [g['gray'][i] for i in range(3)]
    batch = si.detect_and_describe_batch(imgs)
    for i, im in enumerate(imgs):
        (k, d), = si.detect_and_describe_batch([im])
        assert np.array_equal(batch[i][0], k) and np.array_equal(batch[i][1], d)


def test_download_refuses_short_buffer(si, golden):
    """get_all_keypoints copies every image of the batch: a buffer sized for fewer is an error,
    not an overflow."""
    g = golden('parrington')
    counts = si.detect_and_describe_batch([g['gray'][0], g['gray'][1]], download=False)
    with pytest.raises(RuntimeError):
        si.download_results(counts[:1])
    full = si.download_results(counts)
    assert [len(k) for k, _ in full] == [int(c) for c in counts]


def test_empty_and_tiny_images(si):
    kps, desc = si.compute_keypoints_and_descriptors(np.full((40, 48), 7, np.uint8))
    assert kps == [] and desc.shape == (0,) and desc.dtype == np.float32
    kps, desc = si.compute_keypoints_and_descriptors(np.arange(35, dtype=np.uint8).reshape(5, 7))
    assert kps == []


# ----------------------------------------------------------------------------- matcher / vote / projection
@pytest.mark.parametrize('na,nb', [(1, 1), (7, 300), (513, 64), (1601, 578), (2500, 2466), (100, 0), (0, 50)])
def test_matcher_bit_exact_random(iss, oracle, na, nb):
    rng = np.random.default_rng(na * 31 + nb)
    A = rng.integers(0, 256, (na, 128), dtype=np.uint8)
    B = rng.integers(0, 256, (nb, 128), dtype=np.uint8)
    if nb > 10 and na > 0:   # exact ties: duplicated rows -> lowest j must win
        B[nb // 2] = B[3]
        B[nb - 1] = B[3]
        A[0] = B[3]
    idx, d1, d2 = iss.match_descriptors(A, B, return_second=True)
    if na == 0:
        assert len(idx) == 0
        return
    ridx, r1, r2 = oracle.match_u8(A, B)
    assert np.array_equal(idx, ridx) and np.array_equal(d1, r1) and np.array_equal(d2, r2)


def test_match_lists_identical_given_reference_descriptors(iss, si, golden):
    for name in ('out', 'parrington', 'grail'):
        g = golden(name)
        full = set(g['full_images'].tolist())
        for p in range(len(g['n_matches'])):
            if not (p in full and p + 1 in full):
                continue
            ka, kb = si.array_to_keypoints(golden_kps(g, p)), si.array_to_keypoints(golden_kps(g, p + 1))
            ia, ib, matches = iss.match_keypoints(ka, g[f'desc_{p}'], kb, g[f'desc_{p + 1}'].astype(np.float32))
            assert np.array_equal(ia, g[f'match_ia_{p}']) and np.array_equal(ib, g[f'match_ib_{p}'])
            move, pair = iss.ransac(matches, 3)
            report(f'pair {name}[{p}] matches={len(ia)} shift={move} ref={g["shifts"][p]}')
            assert np.array_equal(np.array(move), g['shifts'][p])
            assert np.array_equal(np.array(pair).ravel(), g['best_pairs'][p])
    assert iss.ransac([], 3) == ((0, 0), None)


def test_compute_shift_sift_out_pair(iss, golden):
    g = golden('out')
    move, pair = iss.compute_shift_sift(g['bgr_0'], g['bgr_1'], ransac_thr=3, desc_thresh=25000)
    report(f'compute_shift_sift out: {move} ref {g["shifts"][0]}')
    assert np.abs(np.array(move) - g['shifts'][0]).max() < 0.5
    assert pair is not None


def test_panorama_shifts_parrington_subset(iss, golden):
    g = golden('parrington')
    imgs = [g['gray'][i] for i in range(4)]
    shifts, counts, det = iss.panorama_shifts(imgs, return_details=True)
    for p, s in enumerate(shifts):
        report(f'parrington pair {p}: kps {counts[p]}/{g["n_keypoints"][p]} matches {det[p]["n_matches"]}/'
               f'{g["n_matches"][p]} shift {s} ref {g["shifts"][p]}')
        assert np.abs(np.array(s) - g['shifts'][p]).max() < 0.5


def test_cylindrical_projection_bit_exact(iss, oracle, golden):
    img = natural_image(120, 160, 9, channels=3)
    for f in (704.9, 454.417, 90.0):
        assert np.array_equal(iss.cylindrical_projection(img, f), oracle.cylindrical_projection(img, f))
    # against the unmodified reference's outputs (tests/golden/make_golden_cyl.py)
    g = golden('cyl')
    for i in range(len(g['out_names'])):
        assert np.array_equal(iss.cylindrical_projection(g[f'out_raw_{i}'], float(g['out_focals'][i])), g[f'out_cyl_{i}'])
    for k in range(int(g['n_cases'])):
        assert np.array_equal(iss.cylindrical_projection(g[f'case_img_{k}'], float(g[f'case_focal_{k}'])),
                              g[f'case_out_{k}']), k


def test_pipeline_equals_single_context(si, iss):
    """Throughput mode (two contexts, two host threads on one GPU) returns exactly what the
    one-context path returns, set by set."""
    from vfx_image_stitching_b200.pipeline import PanoramaPipeline
    sets = []
    for k in range(5):
        a = natural_image(160, 200, 20 + k, channels=3)
        sets.append([a, np.roll(a, (2, -15 - k), axis=(0, 1)), np.roll(a, (4, -33), axis=(0, 1))])
    pipe = PanoramaPipeline(depth=2)
    try:
        got = pipe.panorama_shifts(sets)
        assert len(pipe.contexts) == 2
    finally:
        pipe.close()
    for s, (shifts, counts, res) in zip(sets, got):
        ref = si.detect_and_describe_batch(s)
        assert [len(k) for k, _ in ref] == counts.tolist()
        for (k0, d0), (k1, d1) in zip(ref, res):
            assert np.array_equal(k0, k1) and np.array_equal(d0, d1)
        assert shifts == iss.panorama_shifts(s)
