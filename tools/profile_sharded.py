"""Phase wall-clock of panorama.sharded_panorama_shifts under torchrun (debug aid):
torchrun --nproc-per-node 2 tools/profile_sharded.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfx_image_stitching_b200 import _capi, panorama  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'parrington.npz'))['gray']
base = [torch.from_numpy(np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2))).to(dev) for im in g]
imgs = [base[i % 18] for i in range(18 * world)]
ctx = _capi.default_context(local)
backend = panorama.GpuBackend(ctx)
marks = []
panorama._prof = lambda name: (torch.cuda.synchronize(), marks.append((name, time.perf_counter())))
for it in range(6):
    marks.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    panorama.sharded_panorama_shifts(imgs, backend, dist=dist, device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rank == 0 and it >= 3:
        prev = t0
        out = []
        for name, t in marks:
            out.append(f'{name} {1e3 * (t - prev):.3f}')
            prev = t
        print(f'total {1e3 * (t1 - t0):.3f} ms | ' + ' | '.join(out), flush=True)
dist.destroy_process_group()
