"""Round-2 parity tests of the CUDA path (through the C ABI): whole image sets against the reference
goldens, the matcher at sizes where the shared-memory ring and the TMEM double buffer wrap, the
per-call stage functions against every recorded reference call, the ratio test, general descriptor
layouts, a full 4096x3072 frame.

Tolerances (north_star): >= 99 % of the reference's keypoints matched within 0.5 px / 0.05 octave,
descriptor RMS relative L2 < 1e-3, match lists bit-identical given identical descriptors, voted
shift within 0.5 px (asserted here at 1e-3 px).  Integer / index work is bit-exact.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import ROOT, golden_kps, match_keypoint_sets, natural_image
from test_gpu_parity import report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def si():
    from vfx_image_stitching_b200 import sift_impl
    return sift_impl


@pytest.fixture(scope='module')
def iss():
    from vfx_image_stitching_b200 import image_stitching_sift
    return image_stitching_sift


# ----------------------------------------------------------------------------- whole sets vs the reference
@pytest.mark.parametrize('name', ['parrington', 'grail'])
def test_full_set_counts_matches_shifts(si, iss, golden, name):
    """All 18 images and all 17 adjacent pairs of the set (image_stitching_sift.py:312-327) against what
    the unmodified reference produced: keypoint counts, match counts, voted shifts; on the images whose
    keypoints are stored in full also the north_star keypoint / descriptor criteria."""
    g = golden(name)
    imgs = [g['gray'][i] for i in range(len(g['gray']))]
    shifts, counts, det = iss.panorama_shifts(imgs, return_details=True)
    assert len(shifts) == len(imgs) - 1 == len(g['n_matches'])
    dk = np.asarray(counts, np.int64) - g['n_keypoints']
    dm = np.array([d['n_matches'] for d in det], np.int64) - g['n_matches']
    ds = np.abs(np.array(shifts) - g['shifts'])
    report(f'full set {name}: keypoint count diff {dk.tolist()} match count diff {dm.tolist()} '
           f'max |shift - ref| {ds.max():.2e} px')
    assert np.abs(dk).max() <= 2
    assert np.abs(dm).max() <= 1
    assert ds.max() < 1e-3
    res = si.download_results(counts)
    for i in g['full_images'].tolist():
        kps, desc = res[i]
        ref = golden_kps(g, i)
        frac, m = match_keypoint_sets(ref, kps)
        a = g[f'desc_{i}'][m >= 0].astype(np.float64)
        b = desc[m[m >= 0]].astype(np.float64)
        rel = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(a, axis=1), 1)
        rms = float(np.sqrt(np.mean(rel ** 2)))
        report(f'   {name}[{i}] matched {frac:.4f} desc rms rel-L2 {rms:.2e} rows identical {np.mean(rel == 0):.3f}')
        assert frac >= 0.99 and rms < 1e-3


def test_out_pair_counts_and_matches(si, iss, golden):
    g = golden('out')
    shifts, counts, det = iss.panorama_shifts([g['bgr_0'], g['bgr_1']], return_details=True)
    report(f'out pair: keypoints {list(counts)}/{g["n_keypoints"].tolist()} matches {det[0]["n_matches"]}/'
           f'{int(g["n_matches"][0])} shift {shifts[0]} ref {g["shifts"][0]}')
    assert np.abs(np.asarray(counts) - g['n_keypoints']).max() <= 2
    assert abs(det[0]['n_matches'] - int(g['n_matches'][0])) <= 1
    assert np.abs(np.array(shifts[0]) - g['shifts'][0]).max() < 1e-3
    # same descriptors in -> the reference's match list out, index for index
    if list(counts) == g['n_keypoints'].tolist() and det[0]['n_matches'] == int(g['n_matches'][0]):
        assert np.array_equal(det[0]['ia'], g['match_ia_0']) and np.array_equal(det[0]['ib'], g['match_ib_0'])


# ----------------------------------------------------------------------------- matcher where the rings wrap
def _oracle_match(oracle, A, B, threads=None):
    """oracle.match_u8 over row blocks of A on all host threads (ctypes releases the GIL)."""
    threads = threads or os.cpu_count() or 1
    step = max(64, (len(A) + 4 * threads - 1) // (4 * threads))
    blocks = [(i, min(len(A), i + step)) for i in range(0, len(A), step)]
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda b: oracle.match_u8(A[b[0]:b[1]], B), blocks))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))


def _descriptor_sets(golden, kind, na, nb):
    """The three distributions of SURVEY 8(d) config 5 (synthetic.descriptor_sets, shared with the bench)."""
    from vfx_image_stitching_b200.synthetic import descriptor_sets
    g = golden('parrington')
    pool = np.concatenate([g[f'desc_{i}'] for i in g['full_images'].tolist()])
    return descriptor_sets(kind, na, nb, pool)


@pytest.mark.parametrize('na,nb', [(16384, 16384), (2048, 65536)])
@pytest.mark.parametrize('kind', ['real', 'uniform', 'ties'])
def test_matcher_bit_exact_large(iss, oracle, golden, kind, na, nb):
    from vfx_image_stitching_b200 import _capi
    import ctypes as C
    A, B = _descriptor_sets(golden, kind, na, nb)
    idx, d1, d2 = iss.match_descriptors(A, B, return_second=True)
    ctx = _capi.default_context()
    tpc, nch = C.c_int32(), C.c_int32()
    _capi.check(ctx.lib.b200sift_match_grid(ctx.handle, C.byref(tpc), C.byref(nch)))
    ridx, r1, r2 = _oracle_match(oracle, A, B)
    report(f'matcher {kind} {na}x{nb}: tiles/chunk {tpc.value} chunks {nch.value} identical idx '
           f'{np.mean(idx == ridx):.6f} d1 {np.mean(d1 == r1):.6f} d2 {np.mean(d2 == r2):.6f}')
    assert tpc.value >= 3            # >= 3 B tiles per CTA: the 3-stage ring and both TMEM buffers were reused
    assert np.array_equal(idx, ridx) and np.array_equal(d1, r1) and np.array_equal(d2, r2)
    # nearest-only epilogue (the one compute_shift_sift uses) through the ratio entry point's sibling
    idx1, d11 = iss.match_descriptors(A, B)
    assert np.array_equal(idx1, ridx) and np.array_equal(d11, r1)


def test_ratio_test_matches_exact(iss, oracle, golden):
    """sift_visualizeUI.py:247-257 with exact neighbours: 100 d1 < 49 d2 on squared integer distances."""
    g = golden('parrington')
    A, B = g['desc_0'], g['desc_1']
    ia, ib, d1, d2 = iss.ratio_test_matches(A, B, 0.7, return_distances=True)
    ridx, r1, r2 = oracle.match_u8(A, B)
    assert np.array_equal(d1, r1) and np.array_equal(d2, r2)
    keep = 100 * r1.astype(np.int64) < 49 * r2.astype(np.int64)
    assert np.array_equal(ia, np.nonzero(keep)[0]) and np.array_equal(ib, ridx[keep])
    assert 0 < len(ia) < len(A)
    good = iss.good_matches(A, B)
    assert [m.queryIdx for m in good] == ia.tolist() and [m.trainIdx for m in good] == ib.tolist()
    assert abs(good[0].distance - float(np.sqrt(np.float32(r1[ia[0]])))) == 0
    # other ratios, degenerate sets
    ia2, _ = iss.ratio_test_matches(A, B, 0.8)
    assert np.array_equal(ia2, np.nonzero(25 * r1.astype(np.int64) < 16 * r2.astype(np.int64))[0])
    assert len(iss.ratio_test_matches(A, B[:1])[0]) == 0            # no second neighbour -> nothing passes
    assert len(iss.ratio_test_matches(A[:0], B)[0]) == 0


# ----------------------------------------------------------------------------- per-call stage functions
@pytest.fixture(scope='module')
def cv2_pyramid(oracle, golden):
    """Gaussian + DoG pyramid of grail[0] built with the blur the reference calls (cv2.GaussianBlur):
    the exact arrays the reference's own localize / orientation calls saw."""
    pytest.importorskip('cv2')
    g = golden('grail')
    gray = g['gray'][0].astype(np.float32)
    base = oracle.generate_base_image(gray, 1.6, 0.5, 'cv2')
    pyr = oracle.generate_gaussian_images(base, oracle.compute_number_of_octaves(base.shape),
                                          oracle.generate_gaussian_kernels(1.6, 3), 'cv2')
    dog = [[b - a for a, b in zip(o, o[1:])] for o in pyr]
    return g, pyr, dog


def test_localize_matches_every_reference_call(si, cv2_pyramid):
    """loc_0 holds every call of localize_extremum_via_quadratic_fit the reference made on grail[0]
    (sift_impl.py:128-129): same candidates in, same accept / reject and keypoint fields out."""
    g, pyr, dog = cv2_pyramid
    cand, loc = g['cand_0'], g['loc_0']
    for is_dog, layers in ((True, dog), (False, pyr)):
        kps, lyr = si.localize_extrema(cand, layers, is_dog=is_dog)
        ref_ok = loc[:, 6] >= 0
        same_decision = (lyr >= 0) == ref_ok
        both = (lyr >= 0) & ref_ok
        ident = np.ones(len(cand), bool)
        for k, f in enumerate(('x', 'y', 'size', 'response')):
            ident &= kps[f] == loc[:, k].astype(np.float32)
        ident &= kps['octave'] == loc[:, 4].astype(np.int64)
        ident &= lyr == loc[:, 6].astype(np.int64)
        report(f'localize (is_dog={is_dog}) calls {len(cand)} accepted ref {int(ref_ok.sum())} gpu {int((lyr >= 0).sum())} '
               f'same decision {same_decision.mean():.5f} identical fields {ident[both].mean():.5f}')
        assert same_decision.mean() >= 0.999
        assert ident[both].mean() >= 0.995          # float64 adjugate solve vs LAPACK gelsd: last bit of a few
        assert np.abs(kps['x'][both] - loc[both, 0]).max() < 1e-3
    # the per-call form with the reference's own argument list
    n_some = 0
    for j in list(range(0, len(cand), max(1, len(cand) // 25)))[:25]:
        o, l, y, x = (int(v) for v in cand[j])
        res = si.localize_extremum_via_quadratic_fit(x, y, l, o, 3, dog[o], 1.6, 0.04, 5)
        if loc[j, 6] < 0:
            assert res is None
            continue
        kp, lyr1 = res
        assert lyr1 == int(loc[j, 6]) and abs(kp.pt[0] - loc[j, 0]) < 1e-3 and kp.octave == int(loc[j, 4])
        assert kp.angle == -1.0
        n_some += 1
    assert n_some > 3


def test_orientation_counts_match_reference(si, oracle, cv2_pyramid):
    """norient_0: number of keypoints compute_keypoints_with_orientations returned for every localized
    extremum of grail[0] (sift_impl.py:135-137), and the angles against the oracle."""
    g, pyr, dog = cv2_pyramid
    cand, loc, nori = g['cand_0'], g['loc_0'], g['norient_0']
    acc = np.nonzero(loc[:, 6] >= 0)[0]
    assert len(acc) == len(nori)
    total_same, total = 0, 0
    for o in range(len(pyr)):
        for lyr in (1, 2, 3):
            sel = acc[(cand[acc, 0] == o) & (loc[acc, 6] == lyr)]
            if len(sel) == 0:
                continue
            kin = np.zeros(len(sel), si.KP_DTYPE)
            kin['x'], kin['y'], kin['size'], kin['response'] = loc[sel, 0], loc[sel, 1], loc[sel, 2], loc[sel, 3]
            kin['octave'] = loc[sel, 4].astype(np.int64)
            kin['angle'] = -1
            out, counts = si.keypoints_with_orientations(kin, o, pyr[o][lyr])
            ref_counts = nori[np.searchsorted(acc, sel)]
            total_same += int((counts == ref_counts).sum())
            total += len(sel)
            # against the oracle, keypoint by keypoint (angles within 1e-3 deg where the counts agree)
            off = 0
            for i in range(min(len(sel), 40)):
                ro = oracle.compute_keypoints_with_orientations(kin[i], o, pyr[o][lyr])
                got = out[int(counts[:i].sum()):int(counts[:i].sum()) + counts[i]]
                if len(ro) == len(got):
                    d = np.abs(ro['angle'] - got['angle'])
                    assert np.minimum(d, 360 - d).max() < 1e-2
                    assert np.array_equal(ro['x'], got['x']) and np.array_equal(ro['octave'], got['octave'])
                off += counts[i]
    report(f'orientation counts identical for {total_same}/{total} localized keypoints')
    assert total == len(acc) and total_same / total >= 0.995
    # per-call form
    j = acc[0]
    kp, lyr = si.localize_extremum_via_quadratic_fit(int(cand[j, 3]), int(cand[j, 2]), int(cand[j, 1]), int(cand[j, 0]),
                                                     3, dog[cand[j, 0]], 1.6, 0.04, 5)
    kl = si.compute_keypoints_with_orientations(kp, int(cand[j, 0]), pyr[cand[j, 0]][lyr])
    assert len(kl) == nori[0] and kl[0].pt == kp.pt and 0 <= kl[0].angle < 360


def test_find_extrema_reads_caller_dog_images(si, oracle, cv2_pyramid):
    """find_scale_space_extrema(gaussian_images, dog_images, ...): same result as with the on-the-fly
    DoG when dog_images is the true difference, and the caller's DoG is really what is scanned."""
    g, pyr, dog = cv2_pyramid
    a = si.find_scale_space_extrema_array(pyr)
    b = si.find_scale_space_extrema_array(pyr, dog_images=dog)
    assert len(a) == len(b) > 100
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
    flat = [[np.zeros_like(d) for d in o] for o in dog]
    assert len(si.find_scale_space_extrema_array(pyr, dog_images=flat)) == 0
    kps = si.find_scale_space_extrema(pyr, dog, 3, 1.6, 5)
    assert len(kps) == len(a)


@pytest.mark.parametrize('d,nb', [(4, 8), (2, 8), (3, 12), (6, 8), (4, 36)])
def test_descriptors_general_layout(si, oracle, golden, d, nb):
    """generate_descriptors(window_width, num_bins) (sift_impl.py:361-362) for non-default layouts."""
    gray = golden('out')['gray'][1].astype(np.float32)
    base = oracle.generate_base_image(gray, 1.6, 0.5)
    pyr = oracle.generate_gaussian_images(base, oracle.compute_number_of_octaves(base.shape),
                                          oracle.generate_gaussian_kernels(1.6, 3))
    raw = oracle.find_scale_space_extrema(pyr, None, 3, 1.6, 5)
    kps = oracle.convert_keypoints_to_input_image_size(oracle.remove_duplicate_keypoints(raw))[:400]
    ref = oracle.generate_descriptors(kps, pyr, window_width=d, num_bins=nb)
    got = si.generate_descriptors(si.array_to_keypoints(kps), pyr, window_width=d, num_bins=nb)
    assert got.shape == ref.shape == (len(kps), d * d * nb)
    diff = np.abs(got - ref)
    report(f'descriptors d={d} bins={nb}: identical rows {np.mean(diff.sum(1) == 0):.4f} max|diff| {diff.max():.0f}')
    assert diff.max() <= 1 and np.mean(diff.sum(1) == 0) >= 0.97


def test_ransac_float64_and_large(iss, oracle):
    """ransac() (image_stitching_sift.py:86-111) votes on Python floats: float64 coordinates that are
    not float32-representable, and more matches than one shared-memory tile holds."""
    rng = np.random.default_rng(4)
    for n in (1, 5, 1500, 13000):
        m = rng.normal(0, 40, (n, 4))
        m[: n // 3, 2:] = m[: n // 3, :2] - np.array([17.123456789012, -3.987654321098]) + rng.normal(0, 0.4, (n // 3, 2))
        matches = [((a, b), (c, d)) for a, b, c, d in m.tolist()]
        move, pair = iss.ransac(matches, 3)
        rmove, rpair = oracle.ransac(m, 3)
        assert move == tuple(rmove), n
        assert np.array_equal(np.array(pair).ravel(), np.array(rpair).ravel())


def test_contexts_with_different_sigmas_do_not_interfere(si):
    """Two contexts on one GPU blurring with different sigmas from two threads (per-context tap tables)."""
    from vfx_image_stitching_b200 import _capi
    import ctypes as C
    img = (np.random.default_rng(2).random((300, 520)) * 255).astype(np.float32)
    ctxs = [_capi.default_context(), _capi.Context(0)]
    sig = [1.2262735, 3.0900156]
    ref = [si.gaussian_blur(img, s) for s in sig]

    def run(k):
        out = np.empty_like(img)
        for _ in range(40):
            _capi.check(ctxs[k].lib.b200sift_gaussian_blur(ctxs[k].handle, _capi.ptr(img), img.shape[0], img.shape[1],
                                                           C.c_double(sig[k]), _capi.ptr(out), 0))
            if not np.array_equal(out, ref[k]):
                return False
        return True
    with ThreadPoolExecutor(2) as pool:
        assert all(pool.map(run, range(2)))
    ctxs[1].close()


# ----------------------------------------------------------------------------- a full 4096 x 3072 frame
@pytest.mark.slow
def test_full_frame_against_oracle(si, oracle):
    """BASELINE.json configs[3] shape: one synthetic 4096x3072 frame, GPU vs the C oracle (~1 min of CPU)."""
    from vfx_image_stitching_b200.synthetic import natural_image as nat
    frame = nat(3072, 4096, 1000, channels=3)
    (kps, desc), = si.detect_and_describe_batch([frame])
    ref, rdesc = oracle.compute_keypoints_and_descriptors(frame)
    # grid-bucketed matching (the O(n^2) matcher of conftest is too slow for ~15 k keypoints)
    frac, m = _match_sets_bucketed(ref, kps)
    a = np.asarray(rdesc, np.float64)[m >= 0]
    b = desc[m[m >= 0]].astype(np.float64)
    rel = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(a, axis=1), 1)
    rms = float(np.sqrt(np.mean(rel ** 2)))
    report(f'frame 4096x3072: oracle {len(ref)} gpu {len(kps)} matched {frac:.4f} desc rms rel-L2 {rms:.2e}')
    assert frac >= 0.99 and abs(len(kps) - len(ref)) <= max(3, len(ref) // 200) and rms < 1e-3


def _match_sets_bucketed(ref, test, tol_px=0.5, tol_oct=0.05):
    idx = np.full(len(ref), -1, np.int64)
    cell = {}
    for j, (x, y) in enumerate(zip(test['x'], test['y'])):
        cell.setdefault((int(x) // 4, int(y) // 4), []).append(j)
    for i, r in enumerate(ref):
        cx, cy = int(r['x']) // 4, int(r['y']) // 4
        cand = [j for dx in (-1, 0, 1) for dy in (-1, 0, 1) for j in cell.get((cx + dx, cy + dy), ())]
        best, bd = -1, None
        for j in cand:
            t = test[j]
            d2 = (float(t['x']) - float(r['x'])) ** 2 + (float(t['y']) - float(r['y'])) ** 2
            if d2 <= tol_px ** 2 and (t['octave'] & 0xffff) == (r['octave'] & 0xffff) and \
                    abs(np.log2(float(t['size']) / float(r['size']))) <= tol_oct:
                da = abs(((float(t['angle']) - float(r['angle']) + 180) % 360) - 180)
                key = d2 + 1e-3 * da
                if bd is None or key < bd:
                    best, bd = j, key
        idx[i] = best
    return float((idx >= 0).mean()), idx


def test_descriptors_large_and_clipped_windows(si, oracle, golden):
    """The three enumeration paths of describe_kernel: rows of up to 64 px (bands of 30 rows), up to 128 px
    (bands of 15 rows) and the plain path beyond (the huge keypoints the non-convergence quirk of
    sift_impl.py:169-211 can emit), with windows clipped by the image border -- against the oracle."""
    gray = golden('out')['gray'][1].astype(np.float32)
    base = oracle.generate_base_image(gray, 1.6, 0.5)
    pyr = oracle.generate_gaussian_images(base, oracle.compute_number_of_octaves(base.shape),
                                          oracle.generate_gaussian_kernels(1.6, 3))
    raw = oracle.find_scale_space_extrema(pyr, None, 3, 1.6, 5)
    kps = oracle.convert_keypoints_to_input_image_size(oracle.remove_duplicate_keypoints(raw))
    rng = np.random.default_rng(1)
    sel = kps[rng.permutation(len(kps))[:240]].copy()
    for k, f in enumerate((1.0, 2.2, 4.5, 9.0, 20.0, 60.0)):        # half widths ~20 .. > image diagonal
        sel['size'][k::6] *= f
    ref = oracle.generate_descriptors(sel, pyr)
    got = si.generate_descriptors(si.array_to_keypoints(sel), pyr)
    diff = np.abs(got - ref)
    per = [float(np.mean(diff[k::6].sum(1) == 0)) for k in range(6)]
    report(f'descriptors, window scale x1 / 2.2 / 4.5 / 9 / 20 / 60: identical rows {per} max|diff| {diff.max():.0f}')
    assert diff.max() <= 1 and min(per) >= 0.9


def test_descriptors_near_axis_aligned_angles(si, oracle, golden):
    """Row intervals of describe_kernel when sin or cos of the window rotation is tiny: below 1e-3 the term is
    left to the exact per-pixel predicate, above it the analytic bound is used -- angles on both sides of the
    switch around every multiple of 90 degrees, small and large windows, against the oracle."""
    gray = golden('out')['gray'][1].astype(np.float32)
    base = oracle.generate_base_image(gray, 1.6, 0.5)
    pyr = oracle.generate_gaussian_images(base, oracle.compute_number_of_octaves(base.shape),
                                          oracle.generate_gaussian_kernels(1.6, 3))
    raw = oracle.find_scale_space_extrema(pyr, None, 3, 1.6, 5)
    kps = oracle.convert_keypoints_to_input_image_size(oracle.remove_duplicate_keypoints(raw))
    offs = np.array([0.0, 1e-5, 3e-4, 0.03, 0.056, 0.058, 0.07, 0.5], np.float64)     # 1e-3 rad = 0.0573 deg
    angles = np.concatenate([(a + s * offs) % 360.0 for a in (0.0, 90.0, 180.0, 270.0) for s in (1, -1)])
    rng = np.random.default_rng(5)
    sel = kps[rng.permutation(len(kps))[:len(angles) * 3]].copy()
    sel['angle'] = np.tile(angles, 3).astype(np.float32)
    sel['size'][len(angles):2 * len(angles)] *= 2.5
    sel['size'][2 * len(angles):] *= 5.0
    ref = oracle.generate_descriptors(sel, pyr)
    got = si.generate_descriptors(si.array_to_keypoints(sel), pyr)
    diff = np.abs(got - ref)
    report(f'descriptors at near-axis angles: identical rows {np.mean(diff.sum(1) == 0):.4f} max|diff| {diff.max():.0f}')
    assert diff.max() <= 1 and np.mean(diff.sum(1) == 0) >= 0.97
