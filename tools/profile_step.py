#!/usr/bin/env python
"""One headline step (parrington: detect+describe 18 images, match 17 pairs) for ncu.

    python tools/profile_step.py                       # plain run, must exit 0 first
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/r1_step_full -f python tools/profile_step.py

Warm-up steps run outside the profiled range (cudaProfilerStart/Stop bracket ONE step).
`--frames N` profiles detect+describe of N synthetic 4096x3072 frames instead (HBM-resident regime).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=0)
    a = ap.parse_args()
    import torch
    from vfx_image_stitching_b200 import _capi, sift_impl
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    ctx = _capi.default_context(0)
    dev = torch.device('cuda', 0)
    if a.frames:
        from vfx_image_stitching_b200.synthetic import natural_image
        imgs = [torch.from_numpy(natural_image(3072, 4096, 1000 + f, channels=3)).to(dev) for f in range(a.frames)]
        step = lambda: sift_impl.detect_and_describe_batch(imgs, ctx=ctx, download=False)
    else:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'parrington.npz'))
        imgs = [torch.from_numpy(np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2))).to(dev)
                for im in g['gray']]
        step = lambda: iss.panorama_shifts(imgs, ctx=ctx)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok', str(out)[:120])


if __name__ == '__main__':
    main()
