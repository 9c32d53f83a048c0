#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (the headline one is bench.py):

  grail   configs[2]: grail/ 18-image set (fixture tests/golden/grail.npz), detect+describe+match
  out     configs[0]: out/ 2-image pair (the reference's own CPU-runnable case)
  frames  configs[3]: synthetic N x 4096x3072 RGB frames, detect+describe only (the regime where the
                      dense front-end streams from HBM instead of sitting in L2)

One JSON line per configuration on stdout; `--check` also runs the CPU oracle on the first
image / frame and reports the north_star keypoint agreement.

    python bench_configs.py grail out frames --frames 8 --steps 5 [--check] [--out profiles/x.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def timed(fn, steps, warmup, torch, stream, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    out = None
    for s in range(steps):
        flush.fill_(s & 0xff)
        stream.wait_stream(torch.cuda.current_stream())
        ev[s][0].record(stream)
        out = fn()
        ev[s][1].record(stream)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='*', default=['grail', 'out', 'frames'])
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--out', default=None)
    ap.add_argument('--depth', type=int, default=3, help='frames config: batches in flight (library contexts)')
    a = ap.parse_args()
    import torch
    from vfx_image_stitching_b200 import _capi, sift_impl
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    from vfx_image_stitching_b200.synthetic import natural_image
    ctx = _capi.default_context(0)
    dev = torch.device('cuda', 0)
    stream = torch.cuda.Stream(dev)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = []

    def agreement(img, kps):
        from conftest import match_keypoint_sets
        from oracle import sift_oracle as so
        t0 = time.perf_counter()
        ref, _ = so.compute_keypoints_and_descriptors(img)
        frac, _ = match_keypoint_sets(ref, kps)
        return {'oracle_keypoints': int(len(ref)), 'gpu_keypoints': int(len(kps)), 'matched_frac': frac,
                'oracle_s': time.perf_counter() - t0}

    for name in a.configs:
        if name in ('grail', 'out'):
            g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
            imgs = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g['gray']]
            res_t = [torch.from_numpy(im).to(dev) for im in imgs]
            ms, out = timed(lambda: iss.panorama_shifts(res_t, ctx=ctx, return_details=True), a.steps, a.warmup,
                            torch, stream, flush)
            shifts, counts, det = out
            mpix = sum(im.shape[0] * im.shape[1] for im in imgs) / 1e6
            err = float(np.abs(np.array(shifts) - g['shifts']).max())
            line = {'config': name, 'images': len(imgs), 'shape': list(imgs[0].shape), 'ms_per_step': ms,
                    'mpix_per_s': mpix / (ms / 1e3), 'image_pairs_per_s': (len(imgs) - 1) / (ms / 1e3),
                    'keypoints': [int(c) for c in counts], 'reference_keypoints': g['n_keypoints'].tolist(),
                    'matches': [d['n_matches'] for d in det], 'reference_matches': g['n_matches'].tolist(),
                    'max_abs_shift_error_px_vs_reference': err}
        elif name == 'frames':
            frames = [natural_image(3072, 4096, 1000 + f, channels=3) for f in range(a.frames)]
            res_t = [torch.from_numpy(f).to(dev) for f in frames]
            ms, counts = timed(lambda: sift_impl.detect_and_describe_batch(res_t, ctx=ctx, download=False), a.steps,
                               a.warmup, torch, stream, flush)
            mpix = a.frames * 3072 * 4096 / 1e6
            line = {'config': 'frames', 'frames': a.frames, 'shape': [3072, 4096, 3], 'ms_per_step': ms,
                    'mpix_per_s': mpix / (ms / 1e3), 'keypoints_per_frame': [int(c) for c in counts],
                    'algorithmic_dense_bytes': 403 * 3072 * 4096 * a.frames,
                    'dense_GBps_if_all_time_were_dense': 403 * 3072 * 4096 * a.frames / (ms / 1e3) / 1e9,
                    'stats_frame0': list(sift_impl.stage_stats(0, ctx))}
            if a.check:
                kps, _ = sift_impl.download_results(counts, ctx)[0]
                line['check_frame0'] = agreement(frames[0], kps)
            if a.depth > 1:
                # throughput mode: `steps` batches of these frames with `depth` of them in flight
                # (pipeline.PanoramaPipeline: one library context + host thread per batch in flight)
                from vfx_image_stitching_b200.pipeline import PanoramaPipeline
                pipe = PanoramaPipeline(contexts=[ctx] + [_capi.Context(0) for _ in range(a.depth - 1)])
                job = lambda _, c: sift_impl.detect_and_describe_batch(res_t, ctx=c, download=False)  # noqa: E731
                pipe.map(job, range(2 * a.depth))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                k = max(a.steps, 2 * a.depth)
                pipe.map(job, range(k))
                torch.cuda.synchronize()
                e1.record()
                torch.cuda.synchronize()
                ms_t = e0.elapsed_time(e1) / k
                line['throughput'] = {'depth': a.depth, 'batches': k, 'ms_per_batch': ms_t, 'mpix_per_s': mpix / (ms_t / 1e3)}
                pipe.close()
        else:
            raise SystemExit(f'unknown config {name}')
        lines.append(line)
        print(json.dumps(line), flush=True)
    if a.out:
        json.dump(lines, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
