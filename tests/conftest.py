import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box')
    config.addinivalue_line('markers', 'slow: about a minute of CPU oracle time next to the GPU run')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    """name -> npz of the committed reference outputs (tests/golden/make_golden.py)."""
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + '.npz'))
        return cache[name]
    return load


@pytest.fixture(scope='session')
def oracle():
    from oracle import sift_oracle
    sift_oracle.lib()
    return sift_oracle


def natural_image(h, w, seed, channels=1):
    """Deterministic natural-image-like uint8 test image: blobs + edges + mild noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.full((h, w), 96.0, np.float32)
    n_blobs = max(8, h * w // 600)
    for _ in range(n_blobs):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        sy, sx = rng.uniform(1.5, 12), rng.uniform(1.5, 12)
        amp = rng.uniform(-90, 90)
        th = rng.uniform(0, np.pi)
        a = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        b = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        img += amp * np.exp(-0.5 * ((a / sx) ** 2 + (b / sy) ** 2))
    for _ in range(6):
        x0 = int(rng.integers(0, w)); y0 = int(rng.integers(0, h))
        img[y0:, x0:] += rng.uniform(-25, 25)
    img += rng.normal(0, 2.0, (h, w)).astype(np.float32)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if channels == 1:
        return g
    out = np.stack([np.clip(g.astype(np.int32) + d, 0, 255).astype(np.uint8)
                    for d in (rng.integers(-20, 20, (h, w)), 0, rng.integers(-20, 20, (h, w)))], axis=2)
    return np.ascontiguousarray(out)


def match_keypoint_sets(ref, test, tol_px=0.5, tol_oct=0.05):
    """north_star criterion: fraction of ref keypoints with a test keypoint within tol_px, the same
    octave/layer bytes and size within tol_oct octaves.  Returns (fraction, index of the match or -1)."""
    idx = np.full(len(ref), -1, np.int64)
    if len(ref) == 0:
        return 1.0, idx
    if len(test) == 0:
        return 0.0, idx
    tx, ty = test['x'].astype(np.float64), test['y'].astype(np.float64)
    for i, r in enumerate(ref):
        d2 = (tx - r['x']) ** 2 + (ty - r['y']) ** 2
        ok = (d2 <= tol_px ** 2) & ((test['octave'] & 0xffff) == (r['octave'] & 0xffff)) & \
             (np.abs(np.log2(test['size'].astype(np.float64) / r['size'])) <= tol_oct)
        if ok.any():
            da = np.abs(((test['angle'].astype(np.float64) - r['angle'] + 180) % 360) - 180)
            cand = np.where(ok)[0]
            idx[i] = cand[np.argmin(d2[cand] + 1e-3 * da[cand])]
    return float((idx >= 0).mean()), idx


def golden_kps(g, i):
    from oracle.sift_oracle import KP_DTYPE
    kf = g[f'kp_f_{i}']
    out = np.zeros(len(kf), KP_DTYPE)
    out['x'], out['y'], out['size'], out['angle'], out['response'] = kf.T
    out['octave'] = g[f'kp_oct_{i}']
    return out
