"""Multi-GPU sharding of the first loop of run_panorama (image_stitching_sift.py:312-327).

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).
detect+describe is independent per image and matching is independent per adjacent pair (SURVEY
8e), so images are split into contiguous blocks in pano order and pair (i, i+1) belongs to the
rank that owns image i.  The only data-path collective is ONE all-gather of every rank's FIRST
image's descriptors (uint8) + keypoint coordinates: the right-hand side of the previous rank's
boundary pair.  Results are identical for any world size because every image and every pair is
computed by exactly one rank with the same kernels.

The compute sits behind a small backend interface so that the plumbing can be exercised on CPU
with gloo (tests/test_distributed_gloo.py plugs the oracle in); `GpuBackend` is the product.
"""
import ctypes as C

import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: 18 items over 8 ranks -> 3,3,2,2,2,2,2,2."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GpuBackend:
    """CUDA kernels behind the C ABI; results of the local block stay in HBM."""

    def __init__(self, ctx=None):
        from . import _capi
        self.ctx = ctx or _capi.default_context()
        self.counts = np.zeros(0, np.int32)

    def detect(self, images):
        """detect+describe the local block (numpy images or CUDA tensors); returns keypoint counts."""
        from . import sift_impl
        self.counts = np.asarray(sift_impl.detect_and_describe_batch(images, ctx=self.ctx, download=False),
                                 np.int32) if len(images) else np.zeros(0, np.int32)
        return self.counts

    def first_image(self, device):
        """(uint8 (n,128) descriptors, float32 (n,2) xy) of local image 0 as torch tensors on `device`
        (zero-copy views of the context's buffers, copied once into fresh tensors)."""
        import torch
        from ._capi import check
        if len(self.counts) == 0 or self.counts[0] == 0:
            return (torch.zeros((0, 128), dtype=torch.uint8, device=device),
                    torch.zeros((0, 2), dtype=torch.float32, device=device))
        d_desc, d_kps, n = C.c_void_p(), C.c_void_p(), C.c_int32()
        check(self.ctx.lib.b200sift_device_results(self.ctx.handle, 0, C.byref(d_desc), C.byref(d_kps), C.byref(n)))

        class _View:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {'shape': shape, 'typestr': typestr, 'data': (ptr, False),
                                                 'version': 3, 'strides': None}
        desc = torch.as_tensor(_View(d_desc.value, (n.value, 128), '|u1'), device=device).clone()
        kps = torch.as_tensor(_View(d_kps.value, (n.value, 6), '<f4'), device=device)
        return desc, kps[:, :2].contiguous()

    def append_remote(self, desc, xy):
        """Make a gathered (remote) image matchable; returns its local image index."""
        from ._capi import check
        idx = C.c_int32()
        desc = desc.contiguous()
        xy = xy.contiguous()
        check(self.ctx.lib.b200sift_append_results(self.ctx.handle, C.c_void_p(desc.data_ptr()),
                                                   C.c_void_p(xy.data_ptr()), int(desc.shape[0]), 1, C.byref(idx)))
        return idx.value

    def match_pairs(self, pairs, ransac_thr, desc_thresh):
        """[(dx, dy)] for local image-index pairs: matcher + acceptance + vote in one device pass."""
        from . import image_stitching_sift as iss
        return iss.match_pairs(pairs, ransac_thr, desc_thresh, self.ctx)[0]


def sharded_panorama_shifts(images, backend, ransac_thr=3, desc_thresh=25000, dist=None, device='cpu'):
    """All adjacent-pair shifts of `images` (every rank passes the same list; each rank only
    touches its block).  Returns on every rank (shifts [(dx,dy)] * (n-1), keypoint counts [n])."""
    import torch
    n = len(images)
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    counts_local = np.asarray(backend.detect([images[i] for i in range(lo, hi)]), np.int64)

    # ---- exchange: first image of every rank (one padded all-gather of descriptors + xy)
    remote_idx = None
    if world > 1:
        first_d, first_xy = backend.first_image(device)
        cnt = torch.tensor([first_d.shape[0], lo, hi - lo], dtype=torch.int32, device=device)
        cnts = torch.zeros(world * 3, dtype=torch.int32, device=device)   # dim-0 concatenation
        dist.all_gather_into_tensor(cnts, cnt)
        cnts = cnts.cpu().numpy().reshape(world, 3)
        cap = max(1, int(cnts[:, 0].max()))
        # descriptors (128 B) and xy (8 B) of one keypoint travel in one 136-byte row
        row = torch.zeros((cap, 136), dtype=torch.uint8, device=device)
        if first_d.shape[0]:
            row[:first_d.shape[0], :128] = first_d
            row[:first_xy.shape[0], 128:] = first_xy.view(torch.uint8).reshape(-1, 8)
        gathered = torch.zeros((world * cap, 136), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, row)       # the one data-path collective
        gathered = gathered.view(world, cap, 136)
        if hi > lo and hi < n:   # my last image's right neighbour lives on another rank
            src = int(np.nonzero((cnts[:, 1] == hi) & (cnts[:, 2] > 0))[0][0])
            m = int(cnts[src, 0])
            r_desc = gathered[src, :m, :128].contiguous()
            r_xy = gathered[src, :m, 128:].contiguous().view(torch.float32).reshape(m, 2)
            remote_idx = backend.append_remote(r_desc, r_xy)
    else:
        cnts = np.array([[0, lo, hi - lo]])

    # ---- owned pairs, one batched device pass
    pairs, owners = [], []
    for i in range(lo, hi):
        if i + 1 >= n:
            continue
        pairs.append((i - lo, i + 1 - lo) if i + 1 < hi else (i - lo, remote_idx))
        owners.append(i)
    my = np.zeros((max(hi - lo, 0), 3), np.float64)
    if hi > lo:
        my[:, 2] = counts_local
    if pairs:
        for i, s in zip(owners, backend.match_pairs(pairs, ransac_thr, desc_thresh)):
            my[i - lo, 0], my[i - lo, 1] = s

    # ---- results to every rank (tiny): per-image (dx, dy, keypoint count)
    if world > 1:
        maxb = max(1, int(cnts[:, 2].max()))
        buf = torch.zeros((maxb, 3), dtype=torch.float64, device=device)
        if hi > lo:
            buf[:hi - lo] = torch.from_numpy(my).to(device)
        out = torch.zeros((world * maxb, 3), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(out, buf)
        out = out.cpu().numpy().reshape(world, maxb, 3)
        rows = np.concatenate([out[r, :int(cnts[r, 2])] for r in range(world)], 0)
    else:
        rows = my
    shifts = [(float(rows[i, 0]), float(rows[i, 1])) for i in range(n - 1)]
    return shifts, rows[:, 2].astype(np.int64).tolist()
