#!/usr/bin/env python
"""Ceiling probe for the weak-scaling numbers: N ranks on one box each run the SINGLE-GPU headline step
(18 parrington images + 17 pairs, pipeline.PanoramaPipeline, no exchange, no collective inside the timed
region) at the same time.  What this gives per step is what the box -- host cores, PCIe, power -- lets N
independent replicas do; the sharded path of bench.py can only be compared with that, not with N times a
lone GPU.

    python -m torch.distributed.run --nproc-per-node N tools/replicas_probe.py [--depth 4] [--e2e]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--depth', type=int, default=4)
    ap.add_argument('--steps', type=int, default=20)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from vfx_image_stitching_b200 import _capi, sift_impl
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    from vfx_image_stitching_b200.pipeline import PanoramaPipeline
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        sys.stdout.flush()
        fd = os.dup(1); os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev); dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(fd, 1); os.close(fd)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'parrington.npz'))['gray']
    imgs = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g]
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    resident = [t.to(dev) for t in pinned]
    pinned_np = [t.numpy() for t in pinned]
    ctxs = [_capi.default_context(local)] + [_capi.Context(local) for _ in range(a.depth - 1)]
    pipe = PanoramaPipeline(contexts=ctxs)
    pairs = [(i, i + 1) for i in range(len(imgs) - 1)]

    def job_resident(_, c):
        return iss.panorama_shifts(resident, ctx=c)

    def job_e2e(_, c):
        counts = sift_impl.detect_and_describe_batch(pinned_np, ctx=c, download=False)
        shifts = iss.match_pairs(pairs, 3, 25000, c)[0]
        return shifts, sift_impl.download_results(counts, c)

    out = {}
    for name, job in (('value', job_resident), ('e2e', job_e2e)):
        pipe.map(job, range(2 * a.depth))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.map(job, range(a.steps))
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
        tmax, tmin = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        out[name] = {'ms_per_step_max_over_ranks': float(tmax[0]), 'ms_per_step_min_over_ranks': float(tmin[0])}
    if rank == 0:
        print(json.dumps({'replicas': world, 'depth': a.depth, 'host_cpus': os.cpu_count(), **out}))
    pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
