"""Multi-GPU sharding of the first loop of run_panorama (image_stitching_sift.py:312-327).

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).
detect+describe is independent per image and matching is independent per adjacent pair (SURVEY
8e), so images are split into contiguous blocks in pano order and pair (i, i+1) belongs to the
rank that owns image i.  The only exchange is an all-gather of every rank's FIRST image's
descriptors + keypoint coordinates (the right-hand side of the previous rank's boundary pair).
Results are identical for any world size because every image and every pair is computed by
exactly one rank with the same kernels.

The compute is injected (`Ops`) so that the plumbing can be exercised on CPU with gloo.
"""
from dataclasses import dataclass
from typing import Callable, List, Sequence

import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: 18 items over 8 ranks -> 3,3,2,2,2,2,2,2."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class Ops:
    """detect(images) -> list of (kps structured array, uint8 (N,128) descriptors);
    match(kpsA, descA, kpsB, descB, thresh) -> float64 (n,4) match list (xA,yA,xB,yB) in A order;
    vote(matches, thr) -> (dx, dy)."""
    detect: Callable
    match: Callable
    vote: Callable


def gpu_ops(ctx=None):
    """The product ops: CUDA kernels behind the C ABI."""
    from . import image_stitching_sift as iss
    from . import sift_impl

    def detect(images):
        return sift_impl.detect_and_describe_batch(images, ctx=ctx) if len(images) else []

    def match(kA, dA, kB, dB, thresh):
        idx, d2 = iss.match_descriptors(dA, dB, ctx=ctx)
        keep = (d2 < thresh) & (idx != -1)
        ia = np.nonzero(keep)[0]
        ib = idx[keep]
        return np.stack([kA['x'][ia], kA['y'][ia], kB['x'][ib], kB['y'][ib]], 1).astype(np.float64) \
            if len(ia) else np.zeros((0, 4))

    def vote(m, thr):
        if len(m) == 0:
            return (0, 0)
        mv, _ = iss.ransac([((r[0], r[1]), (r[2], r[3])) for r in m], thr, ctx=ctx)
        return mv
    return Ops(detect, match, vote)


def sharded_panorama_shifts(images: Sequence[np.ndarray], ops: Ops, ransac_thr=3, desc_thresh=25000,
                            dist=None, device='cpu'):
    """All adjacent-pair shifts of `images` (every rank passes the same list; each rank only
    touches its block).  Returns on every rank (shifts [(dx,dy)] * (n-1), keypoint counts [n])."""
    import torch
    n = len(images)
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    mine = ops.detect([images[i] for i in range(lo, hi)])

    # ---- exchange: first image of every rank (all-gather; padded to the global max count)
    first_k = mine[0][0] if mine else np.zeros(0, dtype=[('x', 'f4'), ('y', 'f4')])
    first_d = mine[0][1] if mine else np.zeros((0, 128), np.uint8)
    cnt = torch.tensor([len(first_d), lo, hi - lo], dtype=torch.int32, device=device)
    if world > 1:
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        cnts = torch.stack(cnts).cpu().numpy()
        cap = max(1, int(cnts[:, 0].max()))
        d_pad = torch.zeros((cap, 128), dtype=torch.uint8, device=device)
        xy_pad = torch.zeros((cap, 2), dtype=torch.float32, device=device)
        if len(first_d):
            d_pad[:len(first_d)] = torch.from_numpy(np.ascontiguousarray(first_d)).to(device)
            xy_pad[:len(first_d)] = torch.from_numpy(
                np.stack([first_k['x'], first_k['y']], 1).astype(np.float32)).to(device)
        d_all = [torch.zeros_like(d_pad) for _ in range(world)]
        xy_all = [torch.zeros_like(xy_pad) for _ in range(world)]
        dist.all_gather(d_all, d_pad)      # the one data-path collective (descriptors, uint8)
        dist.all_gather(xy_all, xy_pad)
    else:
        cnts = cnt.cpu().numpy()[None]

    # ---- owned pairs
    my_shifts = np.zeros((max(hi - lo, 0), 2), np.float64)
    for i in range(lo, hi):
        if i + 1 >= n:
            continue
        kA, dA = mine[i - lo]
        if i + 1 < hi:
            kB, dB = mine[i + 1 - lo]
        else:  # boundary pair: image i+1 is the first image of the rank whose block starts there
            src = int(np.nonzero((cnts[:, 1] == i + 1) & (cnts[:, 2] > 0))[0][0])
            m = int(cnts[src, 0])
            dB = d_all[src][:m].cpu().numpy()
            xy = xy_all[src][:m].cpu().numpy()
            kB = np.zeros(m, dtype=[('x', 'f4'), ('y', 'f4')])
            kB['x'], kB['y'] = xy[:, 0], xy[:, 1]
        matches = ops.match(kA, dA, kB, dB, desc_thresh)
        my_shifts[i - lo] = ops.vote(matches, ransac_thr)

    # ---- results to every rank (tiny): per-image shift row + keypoint count
    counts_local = np.array([len(k) for k, _ in mine], np.int64)
    if world > 1:
        maxb = int(cnts[:, 2].max())
        buf = torch.zeros((maxb, 3), dtype=torch.float64, device=device)
        if hi > lo:
            buf[:hi - lo, :2] = torch.from_numpy(my_shifts).to(device)
            buf[:hi - lo, 2] = torch.from_numpy(counts_local.astype(np.float64)).to(device)
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        rows = np.concatenate([out[r][:int(cnts[r, 2])].cpu().numpy() for r in range(world)], 0)
    else:
        rows = np.concatenate([my_shifts, counts_local[:, None].astype(np.float64)], 1)
    shifts = [(float(rows[i, 0]), float(rows[i, 1])) for i in range(n - 1)]
    return shifts, rows[:, 2].astype(np.int64).tolist()
