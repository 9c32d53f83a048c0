"""N = 2 on real GPUs (NCCL): the sharded first loop of run_panorama must return exactly what one
process returns.  Needs two visible GPUs; skipped otherwise (the CPU/gloo twin of this test is
tests/test_distributed_gloo.py)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _images(n=7):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'parrington.npz'))['gray'][:n]
    return [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g]


def _worker(rank, world, port, q, min_rows):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from vfx_image_stitching_b200 import _capi, panorama
    if min_rows:
        panorama.MIN_EXCHANGE_ROWS = min_rows   # forces the capacity-growth round of the exchange
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world, device_id=dev)
    try:
        backend = panorama.GpuBackend(_capi.default_context(rank))
        out = None
        for _ in range(2):   # second pass reuses the persistent exchange buffers
            out = panorama.sharded_panorama_shifts(_images(), backend, dist=dist, device=dev)
        # throughput form: 4 jobs through 3 contexts (helper threads for detect, collectives in order)
        backends = [backend, panorama.GpuBackend(_capi.Context(rank)), panorama.GpuBackend(_capi.Context(rank))]
        stream = panorama.sharded_panorama_stream([_images()] * 4, backends, dist=dist, device=dev)
        assert all(s == out[0] and c == out[1] for s, c in stream), rank
        q.put((rank, out[0], out[1]))
        for be in backends:
            be.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('min_rows', [None, 64])
def test_two_gpus_equal_one_process(min_rows):
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    ref_shifts, ref_counts, _ = iss.panorama_shifts(_images(), return_details=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, min_rows)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shifts, counts in got:
        assert counts == [int(c) for c in ref_counts], rank
        assert shifts == [(float(a), float(b)) for a, b in ref_shifts], rank


class _StagedDist:
    """torch.distributed look-alike for a box with ONE GPU: both ranks drive cuda:0 and the collectives
    go over gloo through host staging (gloo moves CPU tensors only).  Everything else -- pack kernel,
    device-counted append, matcher table patch, vote -- is the product's GPU path, unchanged."""

    def __init__(self, dist):
        self._d = dist
        self.isend, self.irecv = 'isend', 'irecv'

    def is_initialized(self):
        return True

    def get_rank(self):
        return self._d.get_rank()

    def get_world_size(self):
        return self._d.get_world_size()

    class P2POp:
        def __init__(self, op, tensor, peer):
            self.op, self.tensor, self.peer = op, tensor, peer

    def batch_isend_irecv(self, ops):
        import torch
        reqs = []
        torch.cuda.current_stream().synchronize()          # the packed rows must exist before they are staged
        for o in ops:
            if o.op == 'isend':
                host = o.tensor.cpu()                       # kept alive by the request below
                reqs.append((self._d.isend(host, o.peer), host, None))
            else:
                host = torch.empty(o.tensor.shape, dtype=o.tensor.dtype)
                reqs.append((self._d.irecv(host, o.peer), host, o.tensor))

        class _Req:
            def __init__(self, r, host, dev):
                self.r, self.host, self.dev = r, host, dev

            def wait(self):
                self.r.wait()
                if self.dev is not None:
                    self.dev.copy_(self.host)
        return [_Req(*r) for r in reqs]

    def all_gather_into_tensor(self, out, inp):
        import torch
        torch.cuda.current_stream().synchronize()
        host_out = torch.empty(out.shape, dtype=out.dtype)
        self._d.all_gather_into_tensor(host_out, inp.cpu())
        out.copy_(host_out)


def _worker_one_gpu(rank, world, port, q, min_rows):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from vfx_image_stitching_b200 import _capi, panorama
    if min_rows:
        panorama.MIN_EXCHANGE_ROWS = min_rows
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        sd = _StagedDist(dist)
        backend = panorama.GpuBackend(_capi.default_context(0))
        out = None
        for _ in range(2):
            out = panorama.sharded_panorama_shifts(_images(), backend, dist=sd, device=dev)
        backends = [backend, panorama.GpuBackend(_capi.Context(0))]
        stream = panorama.sharded_panorama_stream([_images()] * 3, backends, dist=sd, device=dev)
        assert all(s == out[0] and c == out[1] for s, c in stream), rank
        q.put((rank, out[0], out[1]))
        for be in backends:
            be.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,min_rows', [(2, None), (3, 64)])
def test_ranks_sharing_one_gpu_equal_one_process(world, min_rows):
    """The multi-rank GPU path on a ONE-GPU box: `world` processes drive cuda:0, the wire goes over gloo
    with host staging.  Proves block sharding, the packed exchange rows, the device-side count of the
    appended neighbour image (incl. the capacity-growth repeat) and pair ownership with the real kernels:
    the result must equal the single-process one bit for bit."""
    torch = pytest.importorskip('torch')
    import torch.multiprocessing as mp
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    ref_shifts, ref_counts, _ = iss.panorama_shifts(_images(), return_details=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_one_gpu, args=(r, world, port, q, min_rows)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shifts, counts in got:
        assert counts == [int(c) for c in ref_counts], rank
        assert shifts == [(float(a), float(b)) for a, b in ref_shifts], rank
