#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference here.

Test infrastructure only.  Needs /root/reference (sift_impl.py,
image_stitching_sift.py and the image sets); it therefore runs in the build
container only -- the fixtures it writes are what travels to the GPU box.

What is recorded (per image set named on the command line):
  * the SIFT inputs: the cylindrically projected images
    (image_stitching_sift.py:117-136 applied to cv2.imread output), stored as
    the uint8 grey image cv2.cvtColor produces at sift_impl.py:27-28;
  * every call of localize_extremum_via_quadratic_fit (sift_impl.py:169-211):
    arguments (x, y, layer, octave) = the 3x3x3 extrema candidates in scan
    order, and the outcome (None or keypoint fields + final layer);
  * the number of orientations emitted per localized keypoint
    (sift_impl.py:246-293);
  * final keypoints + descriptors of compute_keypoints_and_descriptors
    (sift_impl.py:15-39);
  * per adjacent pair: the literal matcher loop image_stitching_sift.py:63-79
    (run through the real compute_shift_sift with the SIFT call memoised) and
    ransac (:86-111): matched coordinate list, A/B indices, voted shift.

Usage: python tests/golden/make_golden.py [out parrington grail] [--jobs 8]
Writes tests/golden/_full/<set>.npz (git-ignored, complete) and the committed,
size-bounded subsets tests/golden/<set>.npz.
"""
import os
import sys
import argparse
import multiprocessing as mp
import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_set(name):
    sys.path.insert(0, REF)
    import cv2
    import image_stitching_sift as iss
    folder = os.path.join(REF, name)
    paths, focals = iss.read_pano_data(os.path.join(folder, 'pano.txt'))
    names, imgs = [], []
    for p in paths:
        base = p.replace('\\', '/').split('/')[-1]
        img = cv2.imread(os.path.join(folder, base))
        assert img is not None, base
        names.append(os.path.splitext(base)[0])
        imgs.append(img)
    return names, imgs, focals


def _one_image(args):
    """Runs in a worker: projection + instrumented reference SIFT."""
    name, img_bgr, focal = args
    sys.path.insert(0, REF)
    import cv2
    cv2.setNumThreads(1)
    import sift_impl
    import image_stitching_sift as iss
    cyl = iss.cylindrical_projection(img_bgr, focal)
    gray = cv2.cvtColor(cyl, cv2.COLOR_BGR2GRAY)

    cand, loc = [], []
    norient = []
    real_loc = sift_impl.localize_extremum_via_quadratic_fit
    real_ori = sift_impl.compute_keypoints_with_orientations

    def rec_loc(x, y, layer, octave, *a, **k):
        res = real_loc(x, y, layer, octave, *a, **k)
        cand.append((octave, layer, y, x))
        if res is None:
            loc.append((0, 0, 0, 0, 0, 0, -1))
        else:
            kp, lyr = res
            loc.append((kp.pt[0], kp.pt[1], kp.size, kp.response, float(kp.octave), 1.0, lyr))
        return res

    def rec_ori(kp, octave, gimg, *a, **k):
        out = real_ori(kp, octave, gimg, *a, **k)
        norient.append(len(out))
        return out

    sift_impl.localize_extremum_via_quadratic_fit = rec_loc
    sift_impl.compute_keypoints_with_orientations = rec_ori
    try:
        # the CLI hands the projected BGR image to the SIFT entry point
        kps, desc = sift_impl.compute_keypoints_and_descriptors(cyl)
    finally:
        sift_impl.localize_extremum_via_quadratic_fit = real_loc
        sift_impl.compute_keypoints_with_orientations = real_ori
    kp_f = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps],
                    dtype=np.float32).reshape(-1, 5)
    kp_oct = np.array([k.octave for k in kps], dtype=np.int32)
    desc = np.asarray(desc, dtype=np.float32).reshape(-1, 128)
    assert np.all(desc == np.round(desc)) and desc.min(initial=0) >= 0 and desc.max(initial=0) <= 255
    return dict(name=name, gray=gray, bgr=cyl,
                cand=np.array(cand, dtype=np.int32).reshape(-1, 4),
                loc=np.array(loc, dtype=np.float64).reshape(-1, 7),
                norient=np.array(norient, dtype=np.int32),
                kp_f=kp_f, kp_oct=kp_oct, desc=desc.astype(np.uint8))


def _one_pair(args):
    """Literal reference matcher loop + ransac, SIFT memoised."""
    (kfA, dA), (kfB, dB) = args
    sys.path.insert(0, REF)
    import cv2
    import image_stitching_sift as iss

    def mk(kf):
        return [cv2.KeyPoint(float(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4])) for r in kf]
    queue = [(mk(kfA), dA.astype(np.float32)), (mk(kfB), dB.astype(np.float32))]
    captured = {}
    real_ransac = iss.ransac

    def cap_ransac(matches, dist_sq_thresh=3):
        captured['matches'] = list(matches)
        return real_ransac(matches, dist_sq_thresh=dist_sq_thresh)

    iss.compute_keypoints_and_descriptors = lambda img: queue.pop(0)
    iss.ransac = cap_ransac
    move, pair = iss.compute_shift_sift(None, None, ransac_thr=3, desc_thresh=25000)
    m = np.array(captured['matches'], dtype=np.float64).reshape(-1, 4)
    # recover indices (exact float32 coordinates -> first index with that pt, in order)
    ia = []
    ib = []
    # independent exact integer restatement to recover (i, j) -- asserted equal below
    A = dA.astype(np.int64)
    B = dB.astype(np.int64)
    d2 = (A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2 * A @ B.T if len(A) and len(B) else np.zeros((len(A), len(B)), np.int64)
    if d2.size:
        j = d2.argmin(1)
        best = d2[np.arange(len(A)), j]
        keep = best < 25000
        ia = np.nonzero(keep)[0]
        ib = j[keep]
        chk = np.concatenate([kfA[ia, :2], kfB[ib, :2]], axis=1).astype(np.float64)
        assert chk.shape == m.shape and np.array_equal(chk, m), 'integer restatement != reference loop'
    return dict(matches=m, ia=np.asarray(ia, np.int32), ib=np.asarray(ib, np.int32),
                shift=np.array(move, dtype=np.float64),
                pair=np.array(pair if pair is not None else ((0, 0), (0, 0)), dtype=np.float64).reshape(4))


def pack(name, names, res, pairs, keep_images, keep_bgr):
    """keep_images: indices whose full keypoints/descriptors/stage records are stored."""
    out = {'names': np.array(names)}
    out['gray'] = np.stack([r['gray'] for r in res])
    out['n_keypoints'] = np.array([len(r['kp_oct']) for r in res], np.int32)
    out['n_candidates'] = np.array([len(r['cand']) for r in res], np.int32)
    out['n_localized'] = np.array([int((r['loc'][:, 6] >= 0).sum()) for r in res], np.int32)
    out['n_matches'] = np.array([len(p['ia']) for p in pairs], np.int32)
    out['shifts'] = np.stack([p['shift'] for p in pairs]) if pairs else np.zeros((0, 2))
    out['best_pairs'] = np.stack([p['pair'] for p in pairs]) if pairs else np.zeros((0, 4))
    out['full_images'] = np.array(sorted(keep_images), np.int32)
    for i in keep_images:
        r = res[i]
        for k in ('cand', 'loc', 'norient', 'kp_f', 'kp_oct', 'desc'):
            out[f'{k}_{i}'] = r[k]
    for i in keep_bgr:
        out[f'bgr_{i}'] = res[i]['bgr']
    for pi, p in enumerate(pairs):
        if pi in keep_images and pi + 1 in keep_images:
            out[f'match_ia_{pi}'] = p['ia']
            out[f'match_ib_{pi}'] = p['ib']
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('sets', nargs='*', default=['out', 'parrington', 'grail'])
    ap.add_argument('--jobs', type=int, default=os.cpu_count())
    a = ap.parse_args()
    os.makedirs(os.path.join(HERE, '_full'), exist_ok=True)
    import cv2
    meta = f'cv2 {cv2.__version__} ipp={cv2.ipp.useIPP()} numpy {np.__version__}'
    for s in a.sets:
        names, imgs, focals = _load_set(s)
        with mp.Pool(a.jobs) as pool:
            res = pool.map(_one_image, list(zip(names, imgs, focals)), chunksize=1)
            jobs = [((res[i]['kp_f'], res[i]['desc']), (res[i + 1]['kp_f'], res[i + 1]['desc']))
                    for i in range(len(res) - 1)]
            pairs = pool.map(_one_pair, jobs, chunksize=1)
        full = pack(s, names, res, pairs, list(range(len(res))), list(range(len(res))))
        full['meta'] = np.array(meta)
        full['focals'] = np.array(focals)
        np.savez_compressed(os.path.join(HERE, '_full', s + '.npz'), **full)
        n_keep = len(res) if s == 'out' else 3
        small = pack(s, names, res, pairs, list(range(n_keep)), [0, 1] if s == 'out' else [0])
        small['meta'] = np.array(meta)
        small['focals'] = np.array(focals)
        np.savez_compressed(os.path.join(HERE, s + '.npz'), **small)
        print(s, 'images', len(res), 'kps', full['n_keypoints'].tolist(),
              'matches', full['n_matches'].tolist(), flush=True)
        for p in full['shifts']:
            print('   shift', p)


if __name__ == '__main__':
    main()
