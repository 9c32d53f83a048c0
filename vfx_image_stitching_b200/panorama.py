"""Multi-GPU sharding of the first loop of run_panorama (image_stitching_sift.py:312-327).

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).
detect+describe is independent per image and matching is independent per adjacent pair (SURVEY
8e), so images are split into contiguous blocks in pano order and pair (i, i+1) belongs to the
rank that owns image i.  The only data-path exchange is a neighbour transfer: every rank sends its
FIRST image's descriptors (uint8) + keypoint coordinates to the previous rank (NCCL send / recv over
NVLink), the right-hand side of that rank's boundary pair; the voted shifts and keypoint counts are
then all-gathered (24 bytes per image).  Results are identical for any world size because every
image and every pair is computed by exactly one rank with the same kernels.

The compute sits behind a small backend interface so that the plumbing can be exercised on CPU
with gloo (tests/test_distributed_gloo.py plugs the oracle in); `GpuBackend` is the product.
"""
import ctypes as C

import numpy as np

MIN_EXCHANGE_ROWS = 4096   # initial row capacity of the first-image exchange (grown on demand)
_prof = None   # debug hook: tools/profile_sharded.py sets a callable(name)


def _mark(name):
    if _prof is not None:
        _prof(name)


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: 18 items over 8 ranks -> 3,3,2,2,2,2,2,2."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


ROW_BYTES = 136      # wire row of the exchange: 128 descriptor bytes + (x, y) float32
HDR_INTS = 34        # row 0: [keypoints of the image, lo, block length, 0 ...]


class BackendBase:
    """What sharded_panorama_shifts needs from the compute side.

    A backend implements the four primitives detect / first_image / append_remote / match_pairs;
    the three exchange steps below (pack_first, append_exchange, match_pairs_into) are then provided
    with plain tensor operations (this is what the gloo tests run, with the oracle as compute).
    GpuBackend overrides them with single calls into the C library."""

    def stream_ctx(self, device):
        """Context manager under which the exchange (tensor operations + collectives) runs."""
        import contextlib
        return contextlib.nullcontext()

    def close(self):
        """Release the persistent exchange buffers.  Call it before torch.distributed.destroy_process_group()
        and before the library context goes away: the buffers were used by the collectives' streams and on
        the context's stream, and freeing them after either is gone makes the caching allocator touch a
        dead stream."""
        self.__dict__.pop('_xchg', None)

    def pack_first(self, row, cap, tail, device):
        """Fill `row` ((cap+1, 136) uint8 on `device`) with the header + the first local image."""
        import torch
        d, xy = self.first_image(device)
        m = int(d.shape[0])
        hdr = torch.zeros(HDR_INTS, dtype=torch.int32)
        hdr[0] = m
        for i, v in enumerate(tail):
            hdr[1 + i] = int(v)
        row[0].copy_(hdr.view(torch.uint8))
        m = min(m, cap)
        if m:
            row[1:1 + m, :128].copy_(d[:m])
            row[1:1 + m, 128:].copy_(xy[:m].contiguous().view(torch.uint8).view(m, 8))

    def append_exchange(self, recv, cap):
        """Make the image in the wire buffer `recv` ((cap+1, 136) uint8: header row + keypoint rows, as
        pack_first wrote it on the sending rank) matchable; returns its local image index."""
        import torch
        m = min(int(recv[0, :4].cpu().view(torch.int32)[0]), cap)
        r_desc = recv[1:1 + m, :128].contiguous()
        r_xy = recv[1:1 + m, 128:].contiguous().view(torch.float32).view(m, 2)
        return self.append_remote(r_desc, r_xy)

    def match_pairs_into(self, pairs, ransac_thr, desc_thresh, res):
        """Voted (dx, dy) of local pair p into res[p, 0:2] (float64 tensor on the exchange device)."""
        import torch
        if pairs:
            sh = np.asarray(self.match_pairs(pairs, ransac_thr, desc_thresh), np.float64).reshape(-1, 2)
            res[:len(pairs), :2].copy_(torch.from_numpy(sh))


class GpuBackend(BackendBase):
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
        (zero-copy views of the context's buffers)."""
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
        desc = torch.as_tensor(_View(d_desc.value, (n.value, 128), '|u1'), device=device)
        kps = torch.as_tensor(_View(d_kps.value, (n.value, 6), '<f4'), device=device)
        return desc, kps[:, :2]

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

    # ---- exchange steps as single library calls.  The whole exchange -- these kernels, the tensor
    # copies, the neighbour send / recv and the result all-gather -- runs on the CONTEXT's stream
    # (torch.distributed orders its work against the current torch stream), so no events or
    # cross-stream waits are needed and nothing synchronises with the host until _finish().
    def close(self):
        import torch
        BackendBase.close(self)
        self._ext = None
        torch.cuda.synchronize(self.ctx.device)

    def stream_ctx(self, device):
        import torch
        h = self.ctx.stream_handle()
        if getattr(self, '_ext', None) is None or self._ext[0] != h:
            self._ext = (h, torch.cuda.ExternalStream(h, device=device))
        return torch.cuda.stream(self._ext[1])

    def pack_first(self, row, cap, tail, device):
        from ._capi import check
        t = (C.c_int32 * max(1, len(tail)))(*[int(v) for v in tail])
        if len(self.counts) == 0:
            return BackendBase.pack_first(self, row, cap, tail, device)   # empty block: header only
        check(self.ctx.lib.b200sift_pack_exchange(self.ctx.handle, 0, t, len(tail), C.c_void_p(row.data_ptr()), cap))

    def append_exchange(self, recv, cap):
        """No host synchronisation: the count in the header is read on the device (b200sift_append_exchange)."""
        from ._capi import check
        idx = C.c_int32(-1)
        check(self.ctx.lib.b200sift_append_exchange(self.ctx.handle, C.c_void_p(recv.data_ptr()), int(cap),
                                                    C.byref(idx)))
        return idx.value

    def match_pairs_into(self, pairs, ransac_thr, desc_thresh, res):
        from . import image_stitching_sift as iss
        from ._capi import check
        if not pairs:
            return
        pr = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        check(self.ctx.lib.b200sift_match_pairs_device(self.ctx.handle, len(pairs),
                                                       pr.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       iss._int_thresh(desc_thresh),
                                                       float(ransac_thr), C.c_void_p(res.data_ptr()),
                                                       res.stride(0) * res.element_size()))


def sharded_panorama_shifts(images, backend, ransac_thr=3, desc_thresh=25000, dist=None, device='cpu'):
    """All adjacent-pair shifts of `images` (every rank passes the same list; each rank only
    touches its block).  Returns on every rank (shifts [(dx,dy)] * (n-1), keypoint counts [n])."""
    n = len(images)
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    counts_local = np.asarray(backend.detect([images[i] for i in range(lo, hi)]), np.int64)
    _mark('detect')
    while True:
        with backend.stream_ctx(device):
            pending = _exchange_and_match(backend, ransac_thr, desc_thresh, dist, device, rank, world, lo, hi, n,
                                          counts_local)
        res = _finish(pending)
        if res is not None:
            return res
        # a first image did not fit the exchange rows: every rank saw the same counts and grew its buffers


def _exchange_state(backend, world, device):
    st = getattr(backend, '_xchg', None)
    if st is None or st['world'] != world or st['device'] != str(device):
        st = backend._xchg = {'world': world, 'device': str(device), 'cap': 0, 'want_cap': MIN_EXCHANGE_ROWS}
    return st


def _exchange_and_match(backend, ransac_thr, desc_thresh, dist, device, rank, world, lo, hi, n, counts_local):
    """Stage 2 of one image set, ENQUEUED on the backend's stream; returns a handle for _finish().

    Exchange: a rank needs one remote image, the first image of the next rank (right-hand side of its
    block-boundary pair).  Every rank packs its first image into wire rows of 136 bytes (128 B descriptor
    + 8 B xy; row 0 is a header with the keypoint count) and SENDS them to its left neighbour -- a
    point-to-point transfer over NVLink, not an all-gather.  The receiver appends the buffer as an extra
    image whose count is read from the header on the device, so the exchange involves no host
    synchronisation.  Result: per image (dx, dy, keypoint count), one tiny all-gather, copied to pinned
    host memory asynchronously; _finish() waits for it.  If a first image had more keypoints than the wire
    buffer has rows, every rank sees that in the gathered counts, grows the buffer and repeats the set."""
    import torch
    remote_idx = None
    st = None
    if world > 1:
        st = _exchange_state(backend, world, device)
        cap = max(st['want_cap'], MIN_EXCHANGE_ROWS)
        if st['cap'] != cap:
            st['cap'] = cap
            st['row'] = torch.zeros((cap + 1, ROW_BYTES), dtype=torch.uint8, device=device)
            st['recv'] = torch.zeros((cap + 1, ROW_BYTES), dtype=torch.uint8, device=device)
        # blocks are contiguous and empty blocks only occur at the tail, so the owner of image `hi`
        # is simply the next rank, and whoever owns a block sends its first image to the previous rank
        src = rank + 1 if (hi > lo and hi < n) else -1
        dst = rank - 1 if (hi > lo and rank > 0) else -1
        ops = []
        if dst >= 0:
            backend.pack_first(st['row'], cap, (lo, hi - lo), device)
            ops.append(dist.P2POp(dist.isend, st['row'], dst))
        if src >= 0:
            ops.append(dist.P2POp(dist.irecv, st['recv'], src))
        _mark('x.pack')
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()                                   # stream-ordered on CUDA, blocking on CPU (gloo)
        _mark('x.sendrecv')
        if src >= 0:
            remote_idx = backend.append_exchange(st['recv'], cap)
        _mark('exchange')

    # ---- owned pairs in one batched device pass, results to every rank (tiny): per image
    # (dx, dy, keypoint count)
    pairs = []
    for i in range(lo, hi):
        if i + 1 < n:
            pairs.append((i - lo, i + 1 - lo) if i + 1 < hi else (i - lo, remote_idx))
    if world > 1:
        maxb = (n + world - 1) // world                      # largest block (shard_range)
        if st.get('maxb') != maxb:
            st['maxb'] = maxb
            st['res_h'] = torch.zeros((maxb, 3), dtype=torch.float64)
            st['out_h'] = torch.zeros((world * maxb, 3), dtype=torch.float64)
            if torch.device(device).type == 'cuda':
                st['res_h'], st['out_h'] = st['res_h'].pin_memory(), st['out_h'].pin_memory()
            st['res_d'] = torch.zeros((maxb, 3), dtype=torch.float64, device=device)
            st['out_d'] = torch.zeros((world * maxb, 3), dtype=torch.float64, device=device)
        st['res_h'].zero_()
        if hi > lo:
            st['res_h'][:hi - lo, 2] = torch.from_numpy(counts_local.astype(np.float64))
        st['res_d'].copy_(st['res_h'], non_blocking=True)
        backend.match_pairs_into(pairs, ransac_thr, desc_thresh, st['res_d'])
        _mark('match')
        dist.all_gather_into_tensor(st['out_d'], st['res_d'])
        st['out_h'].copy_(st['out_d'], non_blocking=True)
        ev = None
        if torch.device(device).type == 'cuda':
            ev = torch.cuda.Event()
            ev.record()
        return {'st': st, 'event': ev, 'n': n, 'world': world, 'maxb': maxb, 'cap': st['cap']}
    rows = np.zeros((max(hi - lo, 0), 3), np.float64)
    if hi > lo:
        rows[:, 2] = counts_local
    if pairs:
        for p, sft in enumerate(backend.match_pairs(pairs, ransac_thr, desc_thresh)):
            rows[p, 0], rows[p, 1] = sft
    return {'rows': rows, 'n': n}


def _finish(pending):
    """Wait for the results of an enqueued stage 2; returns (shifts, counts), or None when the exchange
    rows were too few for some first image (the wire capacity has then been raised: repeat the set)."""
    n = pending['n']
    if 'rows' in pending:
        rows = pending['rows']
    else:
        st, world, maxb = pending['st'], pending['world'], pending['maxb']
        if pending['event'] is not None:
            pending['event'].synchronize()
        out = st['out_h'].numpy().reshape(world, maxb, 3)
        blocks = [shard_range(n, r, world) for r in range(world)]
        rows = np.concatenate([out[r, :b[1] - b[0]] for r, b in enumerate(blocks)], 0)
        first_counts = [int(out[r, 0, 2]) for r, b in enumerate(blocks) if b[1] > b[0] and r > 0]
        need = max(first_counts) if first_counts else 0
        if need > pending['cap']:
            st['want_cap'] = 1 << (need - 1).bit_length()        # same decision on every rank
            return None
    _mark('results')
    shifts = [(float(rows[i, 0]), float(rows[i, 1])) for i in range(n - 1)]
    return shifts, rows[:, 2].astype(np.int64).tolist()


def sharded_panorama_stream(jobs, backends, ransac_thr=3, desc_thresh=25000, dist=None, device='cpu', after=None):
    """Throughput form of sharded_panorama_shifts for a sequence of image lists (`jobs`).

    Two stages per job: (1) detect+describe of the local block -- no communication, runs on a helper
    thread; (2) exchange + matching + result all-gather -- every collective is issued from the calling
    thread, in job order, so all ranks issue them in the same order.  Stage 2 is only ENQUEUED (no host
    synchronisation inside); its results are collected one job later, after stage 2 of the next job has
    been queued, so the wait and `after(k, backend, shifts, counts)` (optional, e.g. the download of job
    k's keypoints) overlap device work.  Job k uses backends[k % len(backends)] (one library context
    each); a context is handed to a new stage 1 only after its previous job has been collected.  With a
    single backend there is nothing to overlap with: every job runs start to finish on the calling thread.
    Returns [(shifts, counts)] per job -- identical to calling sharded_panorama_shifts job by job."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = list(jobs)
    n_be = len(backends)
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()

    def stage1(k):
        images = jobs[k]
        lo, hi = shard_range(len(images), rank, world)
        return lo, hi, np.asarray(backends[k % n_be].detect([images[i] for i in range(lo, hi)]), np.int64)

    def stage2(k, lo, hi, counts_local):
        be = backends[k % n_be]
        with be.stream_ctx(device):
            return _exchange_and_match(be, ransac_thr, desc_thresh, dist, device, rank, world, lo, hi, len(jobs[k]),
                                       counts_local)

    def collect(k, s1, pending):
        res = _finish(pending)
        while res is None:                         # wire capacity grown: repeat this set's stage 2
            res = _finish(stage2(k, *s1))
        if after is not None:
            after(k, backends[k % n_be], *res)
        return res

    out = [None] * len(jobs)
    lag = min(1, n_be - 1)     # jobs whose collection is deferred behind the next job's stage 2
    with ThreadPoolExecutor(n_be) as pool:
        futs = {k: pool.submit(stage1, k) for k in range(min(n_be, len(jobs)))}
        waiting = []           # (k, stage-1 result, handle) not yet collected
        for k in range(len(jobs)):
            s1 = futs.pop(k).result()
            waiting.append((k, s1, stage2(k, *s1)))
            while len(waiting) > lag:
                j, s1j, pend = waiting.pop(0)
                out[j] = collect(j, s1j, pend)
                if j + n_be < len(jobs):           # the context of job j is free again
                    futs[j + n_be] = pool.submit(stage1, j + n_be)
        for j, s1j, pend in waiting:
            out[j] = collect(j, s1j, pend)
    return out
