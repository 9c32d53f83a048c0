"""The N>1 plumbing (image-block sharding, all-gather of block-first descriptors, pair ownership)
on CPU with gloo, world_size 2 and 3.  Compute is injected: here the oracle plays the kernels, so
the test checks that every world size returns exactly the single-process result."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_backend():
    from vfx_image_stitching_b200.panorama import BackendBase

    class OracleBackend(BackendBase, _OraclePrimitives):
        pass
    return OracleBackend()


class _OraclePrimitives:
    """The backend primitives of panorama.py with the CPU oracle playing the kernels."""

    def __init__(self):
        from oracle import sift_oracle as so
        self.so = so
        self.res = []

    def detect(self, images):
        self.res = []
        for im in images:
            k, d = self.so.compute_keypoints_and_descriptors(im)
            self.res.append((np.stack([k['x'], k['y']], 1).astype(np.float32).reshape(-1, 2),
                             np.asarray(d, np.float32).reshape(-1, 128).astype(np.uint8)))
        return np.array([len(d) for _, d in self.res], np.int32)

    def first_image(self, device):
        import torch
        if not self.res:
            return torch.zeros((0, 128), dtype=torch.uint8), torch.zeros((0, 2), dtype=torch.float32)
        xy, d = self.res[0]
        return torch.from_numpy(d.copy()), torch.from_numpy(xy.copy())

    def append_remote(self, desc, xy):
        self.res.append((xy.numpy().copy(), desc.numpy().copy()))
        return len(self.res) - 1

    def match_pairs(self, pairs, ransac_thr, desc_thresh):
        out = []
        for a, b in pairs:
            (xa, da), (xb, db) = self.res[a], self.res[b]
            idx, d1, _ = self.so.match_u8(da, db)
            keep = (d1 < desc_thresh) & (idx != -1)
            ia = np.nonzero(keep)[0]
            ib = idx[keep]
            m = np.concatenate([xa[ia], xb[ib]], 1).astype(np.float64) if len(ia) else np.zeros((0, 4))
            out.append(tuple(self.so.ransac(m, ransac_thr)[0]))
        return out


def _images():
    from vfx_image_stitching_b200.synthetic import panorama_set
    return panorama_set(5, 96, 128, seed=7, shift=(-2, -40))


def _worker(rank, world, port, q, min_rows=None):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from vfx_image_stitching_b200 import panorama
    from vfx_image_stitching_b200.panorama import sharded_panorama_shifts
    if min_rows:
        panorama.MIN_EXCHANGE_ROWS = min_rows   # forces the capacity-growth round of the exchange
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        shifts, counts = sharded_panorama_shifts(_images(), _oracle_backend(), dist=dist, device='cpu')
        # throughput form: three jobs through two backends must give the same answer every time
        stream = panorama.sharded_panorama_stream([_images()] * 3, [_oracle_backend(), _oracle_backend()], dist=dist,
                                                  device='cpu')
        assert all(s == shifts and c == counts for s, c in stream), rank
        q.put((rank, shifts, counts))
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    from vfx_image_stitching_b200.panorama import shard_range
    assert [shard_range(18, r, 8)[1] - shard_range(18, r, 8)[0] for r in range(8)] == [3, 3, 2, 2, 2, 2, 2, 2]
    for n, w in ((18, 1), (18, 4), (5, 8), (64, 8), (0, 2)):
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))


@pytest.mark.parametrize('world,min_rows', [(2, None), (3, None), (2, 8)])
def test_sharded_equals_single_process(world, min_rows):
    import torch.multiprocessing as mp
    from vfx_image_stitching_b200.panorama import sharded_panorama_shifts
    ref_shifts, ref_counts = sharded_panorama_shifts(_images(), _oracle_backend())
    assert len(ref_shifts) == 4 and sum(abs(abs(s[0]) - 40) < 1.0 and abs(abs(s[1]) - 2) < 1.0 for s in ref_shifts) >= 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, min_rows)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shifts, counts in got:
        assert shifts == ref_shifts and counts == ref_counts, rank


@pytest.mark.parametrize('n_backends', [1, 2, 3])
def test_stream_distinct_jobs_keep_their_own_results(n_backends):
    """sharded_panorama_stream with DISTINCT jobs: job k must get job k's results for any number of
    backends -- with a single backend stage 1 of job k+1 must not start before stage 2 of job k has
    read that backend's results (round-1 advisor finding)."""
    import time
    from vfx_image_stitching_b200.panorama import BackendBase, sharded_panorama_stream

    class Tagged(BackendBase):
        def __init__(self):
            self.tag = None

        def detect(self, images):
            self.tag = images[0]
            return np.array([images[0]] * len(images), np.int32)

        def match_pairs(self, pairs, ransac_thr, desc_thresh):
            time.sleep(0.01)                       # stage 2 reads the backend's state late
            return [(float(self.tag), float(p[1])) for p in pairs]

    jobs = [[k + 1] * 3 for k in range(6)]
    seen = []
    out = sharded_panorama_stream(jobs, [Tagged() for _ in range(n_backends)],
                                  after=lambda k, be, shifts, counts: seen.append((k, be.tag)))
    assert [k for k, _ in seen] == list(range(6))
    for k, (shifts, counts) in enumerate(out):
        assert shifts == [(float(k + 1), 1.0), (float(k + 1), 2.0)], (k, shifts)
        assert counts == [k + 1] * 3
