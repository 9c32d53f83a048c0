"""Deterministic synthetic inputs of the shapes named in BASELINE.json (bench + tests)."""
import numpy as np


def natural_image(h, w, seed, channels=1):
    """Natural-image-like uint8 image: anisotropic Gaussian blobs + step edges + 2-grey-level noise
    (white noise alone would explode the extrema count)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.full((h, w), 96.0, np.float32)
    n_blobs = max(8, h * w // 600)
    for _ in range(n_blobs):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        sy, sx = rng.uniform(1.5, 12), rng.uniform(1.5, 12)
        amp = rng.uniform(-90, 90)
        th = rng.uniform(0, np.pi)
        # evaluate each blob only inside its 4-sigma box (keeps large frames cheap)
        r = int(4 * max(sx, sy)) + 1
        y0, y1 = max(0, int(cy) - r), min(h, int(cy) + r + 1)
        x0, x1 = max(0, int(cx) - r), min(w, int(cx) + r + 1)
        if y0 >= y1 or x0 >= x1:
            continue
        X, Y = xx[y0:y1, x0:x1] - cx, yy[y0:y1, x0:x1] - cy
        a = X * np.cos(th) + Y * np.sin(th)
        b = -X * np.sin(th) + Y * np.cos(th)
        img[y0:y1, x0:x1] += amp * np.exp(-0.5 * ((a / sx) ** 2 + (b / sy) ** 2))
    for _ in range(6):
        x0 = int(rng.integers(0, w))
        y0 = int(rng.integers(0, h))
        img[y0:, x0:] += rng.uniform(-25, 25)
    img += rng.normal(0, 2.0, (h, w)).astype(np.float32)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if channels == 1:
        return g
    return np.ascontiguousarray(np.repeat(g[:, :, None], 3, axis=2))


def panorama_set(n_images=18, h=512, w=384, seed=1000, shift=(-3, -245)):
    """n overlapping views (uint8 BGR, h x w) cut from one wide synthetic scene: consecutive views
    are displaced by `shift` (dy, dx), the way a tripod sequence such as parrington/ is."""
    dy, dx = shift
    H = h + abs(dy) * (n_images - 1) + 8
    W = w + abs(dx) * (n_images - 1) + 8
    scene = natural_image(H, W, seed)
    out = []
    for i in range(n_images):
        y0 = (abs(dy) * (n_images - 1 - i)) if dy < 0 else dy * i
        x0 = (abs(dx) * i) if dx < 0 else dx * (n_images - 1 - i)
        v = scene[y0:y0 + h, x0:x0 + w]
        out.append(np.ascontiguousarray(np.repeat(v[:, :, None], 3, axis=2)))
    return out


def mosaic_frame(frame, tiles, rows=8, cols=8):
    """BASELINE.json configs[3], variant (i) of SURVEY 8(d): a 3072 x 4096 BGR frame tiled 8 x 8 from the
    repo's 384 x 512 images rotated by 90 degrees; tile (r, c) of frame f is image (f + 8 r + c) mod len(tiles).
    `tiles`: list of 2-D uint8 images of identical shape (the cylindrically projected parrington + grail
    images of tests/golden/*.npz, 512 rows x 384 cols each).  Natural-image content: ~1.5 k keypoints per tile."""
    th, tw = tiles[0].shape[1], tiles[0].shape[0]          # after rot90: 384 rows x 512 cols
    out = np.empty((rows * th, cols * tw), np.uint8)
    for r in range(rows):
        for c in range(cols):
            t = np.rot90(tiles[(frame + cols * r + c) % len(tiles)])
            out[r * th:(r + 1) * th, c * tw:(c + 1) * tw] = t
    return np.ascontiguousarray(np.repeat(out[:, :, None], 3, axis=2))


def descriptor_sets(kind, na, nb, pool=None, seed=None):
    """The descriptor distributions of BASELINE.json configs[4]: 'real' -- rows of `pool`
    (reference descriptors) with +-2 jitter, seed 7; 'uniform' -- rng.integers(0, 256), seed 8; 'ties' --
    uniform with 1 % duplicated rows (exact ties: the lowest j must win)."""
    if kind == 'real':
        rng = np.random.default_rng(7 if seed is None else seed)
        pool = np.asarray(pool).astype(np.int16)

        def draw(n):
            rows = pool[rng.integers(0, len(pool), n)]
            return np.clip(rows + rng.integers(-2, 3, rows.shape), 0, 255).astype(np.uint8)
        return draw(na), draw(nb)
    rng = np.random.default_rng(8 if seed is None else seed)
    A = rng.integers(0, 256, (na, 128), dtype=np.uint8)
    B = rng.integers(0, 256, (nb, 128), dtype=np.uint8)
    if kind == 'ties':
        dup = rng.integers(0, nb, max(1, nb // 100))
        B[dup] = B[rng.integers(0, nb, len(dup))]
        A[rng.integers(0, na, max(1, na // 100))] = B[rng.integers(0, nb, max(1, na // 100))]
    return A, B
