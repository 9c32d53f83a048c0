#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200:
SIFT detect+describe Mpix/s & match pairs/s on the parrington configuration
(18 images 384x512, 17 adjacent pairs; configs[1]) at N = 1/2/4/8 GPUs, next to the CPU
restatement of the reference timed on the host cores.

    python bench.py --gpus 1 --steps 20 --warmup 3
    torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)
    python bench.py --impl reference ...                           (CPU arm: the oracle port)

A "step" = one pass of the hot path over the 18-image set: detect+describe of every image
(each image once) + brute-force matching and the translation vote of the 17 adjacent pairs.
At N > 1 GPUs the job is a chain of N such sets (18 N images in pano order, 18 N - 1 adjacent
pairs), sharded in contiguous blocks of 18 images per rank: per-GPU work is fixed ("weak"
scaling), block-boundary pairs need the neighbour rank's first image (one NCCL send / recv).  The
strong-scaling time of the single 18-image set over N ranks is reported next to it
(`strong_18_images`).
`value`  : inputs already resident in HBM; throughput over K steps with --depth steps in flight per GPU
           (one library context per step in flight: the host round trips of one step overlap the kernels
           of the others), device-timed with CUDA events around all K steps.
`e2e`    : the same through the drop-in API with pinned HOST images in, host keypoints /
           descriptors / shifts out (H2D + D2H inside the timed region).
`single_step`: one step at a time on one context (the latency of a single call), per-step events.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'sift_detect_describe_match_mpix_per_s'
UNIT = 'Mpix/s'
# the same string in both arms (the driver compares the two config.workload fields)
WORKLOAD = 'parrington 18 x 384x512 detect+describe + 17 adjacent-pair match + vote (BASELINE.json configs[1])'


def load_workload():
    """18 BGR uint8 images 512 rows x 384 cols.  The reference's own parrington set (cylindrically
    projected, as the CLI feeds SIFT) when the fixture is present, else a synthetic panorama."""
    fx = os.path.join(ROOT, 'tests', 'golden', 'parrington.npz')
    if os.path.exists(fx):
        g = np.load(fx)['gray']
        imgs = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, axis=2)) for im in g]
        return imgs, 'parrington/ 18 x 384x512 (reference images, cylindrically projected; fixture tests/golden/parrington.npz)'
    from vfx_image_stitching_b200.synthetic import panorama_set
    return panorama_set(18, 512, 384), 'synthetic 18 x 384x512 panorama sequence'


def mosaic_tiles():
    """The 36 reference images (parrington + grail fixtures) the synthetic frames are tiled from."""
    tiles = []
    for name in ('parrington', 'grail'):
        fx = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
        if os.path.exists(fx):
            tiles += list(np.load(fx)['gray'])
    if not tiles:
        from vfx_image_stitching_b200.synthetic import natural_image
        tiles = [natural_image(512, 384, 50 + i) for i in range(36)]
    return tiles


def int8_peak(torch, dev):
    """Dense int8 tensor peak measured now on this GPU (cuBLASLt IGEMM 8192^3 through torch._int_mm, best of
    8 launches): the denominator of the matcher's roofline fraction (it issues tcgen05.mma kind::i8)."""
    try:
        n = 8192
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev).t()
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, 'measured in this run (torch._int_mm 8192^3, cuBLASLt IGEMM, best of 8)'
    except Exception as e:   # noqa: BLE001
        p = os.path.join(ROOT, 'profiles', 'r2_int8_peak.json')
        if os.path.exists(p):
            return float(json.load(open(p))['int8_tops']), f'profiles/r2_int8_peak.json (live probe failed: {e!r})'[:160]
        return 2.0 * 1662.8, 'fallback: 2 x measured bf16 peak'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_run(imgs, steps, warmup, budget_s=150.0):
    """The oracle port (oracle/sift_oracle.c -- CPU restatement of the reference, pinned against the
    unmodified Python reference) on all host threads.  Returns (Mpix/s, ms/step, sample text,
    threads, desc-pairs/s)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import sift_oracle as so
    so.lib()
    threads = os.cpu_count() or 1
    pool = ThreadPoolExecutor(threads)

    def one_step(sub):
        res = list(pool.map(lambda im: so.compute_keypoints_and_descriptors(im), sub))   # ctypes drops the GIL
        def pair(i):
            _, _, m = so.match_pairs(res[i][0], res[i][1], res[i + 1][0], res[i + 1][1])
            return so.ransac(m, 3)[0], len(res[i][1]) * len(res[i + 1][1])
        out = list(pool.map(pair, range(len(sub) - 1)))
        return sum(o[1] for o in out)

    t0 = time.perf_counter()
    one_step(imgs[:2])
    per2 = time.perf_counter() - t0
    # bounded sample: as many images of the set as fit the budget
    est_full = per2 / 2 * len(imgs) / min(threads, len(imgs)) * 2.5 + 0.05
    n = len(imgs)
    while n > 2 and est_full * n / len(imgs) * (steps + warmup) > budget_s:
        n -= 1
    sub = imgs[:n]
    for _ in range(warmup):
        one_step(sub)
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(steps):
        pairs = one_step(sub)
    dt = (time.perf_counter() - t0) / steps
    mpix = sum(im.shape[0] * im.shape[1] for im in sub) / 1e6
    sample = f'first {n} of 18 images + their {n - 1} adjacent pairs, per step' if n < len(imgs) else \
        'full 18-image set + 17 pairs, per step'
    return mpix / dt, dt * 1e3, sample, threads, pairs / dt


def ncu_traffic(path):
    """{radius: dram read + write bytes per launch} parsed from a committed ncu summary (profiles/*.txt)."""
    import re
    if not os.path.exists(path):
        return None, f'{os.path.basename(path)} missing'
    out, cur = {}, None
    for ln in open(path):
        m = re.match(r'== void blur_ring_kernel<(\d+),', ln)
        if m:
            cur = int(m.group(1))
            out[cur] = 0.0
        m = re.match(r'\s+dram__bytes_(read|write)\.sum\s+(\d+) byte', ln)
        if m and cur is not None:
            out[cur] += float(m.group(2))
    ok = all(r in out for r in (5, 6, 8, 10, 13))
    return (out if ok else None), f'parsed from profiles/{os.path.basename(path)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)'


def reference_arm(args, rank):
    if rank != 0:
        return
    imgs, data = load_workload()
    val, ms, sample, threads, dps = cpu_run(imgs, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': data,
        'config': {'workload': WORKLOAD,
                   'note': 'the reference is pure Python (no compiled sources, oracle/_ref does not exist) and is absent '
                           'from the GPU box; this arm times the C restatement oracle/sift_oracle.c (pinned against the '
                           'reference by tests/golden) on all host threads.  The Python reference itself: 0.005-0.008 '
                           'Mpix/s (BASELINE.md, 8 cores)'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'match_desc_pairs_per_s': dps, 'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--depth', type=int, default=4, help='steps in flight per GPU (library contexts per rank)')
    ap.add_argument('--sync-mode', type=int, default=-1, help='host wait of the library: 0 spin, 1 poll+yield, 2 sleep (-1: library default)')
    ap.add_argument('--no-extra', action='store_true', help='skip the frames (configs[3]) and matcher (configs[4]) legs')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return reference_arm(args, rank)

    import torch
    import torch.distributed as dist
    from vfx_image_stitching_b200 import _capi, panorama, sift_impl
    from vfx_image_stitching_b200 import image_stitching_sift as iss
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: no CUDA device visible (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL_DEBUG stays as the caller set it: the driver reads the communicator's init lines.  NCCL
        # prints its version banner (NCCL_DEBUG=VERSION, the default of this image) and, without
        # NCCL_DEBUG_FILE, its INFO lines to stdout; stdout is the channel of the one JSON line, so file
        # descriptor 1 points at stderr while the communicator is created (eager with device_id) and used once.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    ctx = _capi.default_context(local)
    stream = torch.cuda.Stream(dev, priority=-1)     # the library's side streams run at the lowest priority
    ctx.set_stream(stream.cuda_stream)

    base_imgs, data = load_workload()
    nb = len(base_imgs)
    h, w = base_imgs[0].shape[:2]
    n = nb * world                                   # chain of `world` sets: weak scaling
    mpix_step = n * h * w / 1e6
    pinned_base = [torch.from_numpy(im).pin_memory() for im in base_imgs]   # e2e inputs (host, pinned)
    resident_base = [t.to(dev) for t in pinned_base]                        # value-leg inputs (HBM)
    pinned_np = [pinned_base[i % nb].numpy() for i in range(n)]
    resident = [resident_base[i % nb] for i in range(n)]
    imgs = base_imgs
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    backend = panorama.GpuBackend(ctx)
    lo, hi = panorama.shard_range(n, rank, world)

    def step_resident():
        """inputs in HBM; per-pair results stay on the device except counts / the voted shift."""
        if world == 1:
            return iss.panorama_shifts(resident, ctx=ctx)          # [(dx, dy)] * 17, like the reference loop
        return panorama.sharded_panorama_shifts(resident, backend, dist=dist, device=dev)

    # e2e outputs land in pinned host memory as well (allocated once, grown on demand)
    out_pin = {}

    def pinned_out_for(slot, total):
        st = out_pin.setdefault(slot, {})
        if st.get('cap', 0) < total:
            cap = max(2 * total, 1 << 16)
            kp_bytes = torch.empty(cap * sift_impl.KP_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
            desc = torch.empty((cap, 128), dtype=torch.uint8).pin_memory()
            st.update(cap=cap, keep=(kp_bytes, desc), kps=kp_bytes.numpy().view(sift_impl.KP_DTYPE), desc=desc.numpy())
        return st['kps'], st['desc']

    def e2e_job(_, c):
        """host images in, host keypoints + descriptors + shifts out, on context `c`."""
        counts = sift_impl.detect_and_describe_batch(pinned_np, ctx=c, download=False)
        shifts = iss.match_pairs([(i, i + 1) for i in range(n - 1)], 3, 25000, c)[0]
        res = sift_impl.download_results(counts, c, out=pinned_out_for(id(c), int(np.sum(counts))))
        return shifts, counts, res

    DEPTH = args.depth      # image sets in flight per GPU (library contexts per rank)
    extra_ctx = [_capi.Context(local) for _ in range(DEPTH - 1)]
    all_ctx = [ctx] + extra_ctx
    if args.sync_mode >= 0:
        for c_ in all_ctx:
            c_.set_sync_mode(args.sync_mode)

    def launches_now():
        return sum(c.launch_count() for c in all_ctx)

    def timed_throughput(run, steps, warmup):
        """`run(k)` processes k independent steps with DEPTH of them in flight.  Device time between two
        events bracketing ALL steps (barrier + full device synchronisation on both sides) / steps, max
        over ranks.  The per-step working set (0.47 GB of pyramid per 18 images) exceeds the 126 MB L2, so
        consecutive steps cannot feed each other from cache; no flush is written between them."""
        run(max(warmup, 2 * DEPTH))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = launches_now()
        e0.record()
        t0 = time.perf_counter()
        out = run(steps)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps * 1e3
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out, (launches_now() - l0) / steps

    if world == 1:
        # N = 1: pipeline.PanoramaPipeline -- DEPTH contexts, one host thread each; the uploads / downloads /
        # host round trips of one set overlap the kernels of another
        from vfx_image_stitching_b200.pipeline import PanoramaPipeline
        pipe = PanoramaPipeline(contexts=all_ctx)

        def run_resident(k):
            return pipe.map(lambda _, c_: (iss.panorama_shifts(resident, ctx=c_), None), range(k))[-1]

        def run_e2e(k):
            return pipe.map(e2e_job, range(k))[-1]
    else:
        # N > 1: panorama.sharded_panorama_stream -- detect+describe of step k+1 (helper thread, other
        # context) next to the exchange + matching of step k; all collectives from this thread, in order
        backends = [backend] + [panorama.GpuBackend(c_) for c_ in extra_ctx]

        def run_resident(k):
            return panorama.sharded_panorama_stream([resident] * k, backends, dist=dist, device=dev)[-1]

        def run_e2e(k):
            def after(j, be, shifts, counts):            # this rank's block, to the (pinned) host
                sift_impl.download_results(be.counts, be.ctx, out=pinned_out_for(id(be.ctx), int(np.sum(be.counts))))
            return panorama.sharded_panorama_stream([pinned_np] * k, backends, dist=dist, device=dev, after=after)[-1]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        out = None
        for s in range(steps):
            flush.fill_(s & 0xff)                       # L2 flush between steps (untimed)
            stream.wait_stream(torch.cuda.current_stream(dev))
            ev[s][0].record(stream)
            out = fn()
            ev[s][1].record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps * 1e3
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out, (ctx.launch_count() - l0) / steps

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, wall_dev, out_dev, launches = timed_throughput(run_resident, args.steps, args.warmup)
    ms_e2e, wall_e2e, out_e2e, _ = timed_throughput(run_e2e, args.steps, args.warmup)
    # one step at a time on one context, per-step CUDA events, 256 MiB L2 flush before every step: the
    # latency of a single call, reported next to the throughput numbers
    ms_one, _, _, launches_one = timed(step_resident, max(5, args.steps // 2), 3)
    clocks = sampler.stop() if sampler else None
    strong = None
    if world > 1:   # the single 18-image set split over all ranks (strong scaling), for the record
        def step_strong():
            return panorama.sharded_panorama_shifts(resident_base, backend, dist=dist, device=dev)
        ms_s, _, _, _ = timed(step_strong, max(3, args.steps // 2), 3)
        strong = {'ms_per_step': ms_s, 'value': nb * h * w / 1e6 / (ms_s / 1e3), 'unit': UNIT,
                  'sharding': '18 images in contiguous blocks over all ranks'}

    # ---- BASELINE.json configs[3]: 64 synthetic 4096x3072 frames, detect+describe only, the frames split
    # over the ranks (strong scaling, no exchange); configs[4]: the 64k x 64k matcher, A rows split over the
    # ranks, B replicated (no reduction: the top-2 is per A row).  Both on every rank, max time over ranks.
    import ctypes as C
    frames_line = matcher_line = None
    if not args.no_extra:
        from vfx_image_stitching_b200.synthetic import descriptor_sets, mosaic_frame
        tiles = mosaic_tiles()
        f_lo, f_hi = panorama.shard_range(64, rank, world)
        FB = 8                                            # frames per detect call
        my_frames = [torch.from_numpy(mosaic_frame(f, tiles)).to(dev) for f in range(f_lo, f_hi)]
        batches = [my_frames[i:i + FB] for i in range(0, len(my_frames), FB)]
        from vfx_image_stitching_b200.pipeline import PanoramaPipeline
        fpipe = PanoramaPipeline(contexts=all_ctx)

        def run_frames(reps):
            jobs = [b for _ in range(reps) for b in batches]
            return fpipe.map(lambda b, c_: sift_impl.detect_and_describe_batch(b, ctx=c_, download=False), jobs)
        if batches:
            run_frames(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        FREPS = 2
        e0.record()
        fcounts = run_frames(FREPS) if batches else []
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        kp_local = float(sum(int(np.sum(cn)) for cn in fcounts[:len(batches)]))
        t = torch.tensor([e0.elapsed_time(e1) / FREPS, kp_local], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tmax[0]
        ms_f = float(t[0])
        fpx = 64 * 3072 * 4096
        frames_line = {
            'workload': 'synthetic 64 x 4096x3072 RGB frames, detect+describe (BASELINE.json configs[3]); mosaic of the '
                        'reference images (8 x 8 tiles of 384x512, SURVEY 8d variant i), frames split over the ranks',
            'scaling': 'strong', 'frames_per_rank': -(-64 // world), 'frames_per_call': FB, 'ms': ms_f,
            'value': fpx / 1e6 / (ms_f / 1e3), 'unit': UNIT, 'keypoints': int(float(t[1])),
            'algorithmic_dense_bytes': 403 * fpx,
            'dense_GBps_if_all_time_were_dense': 403 * fpx / (ms_f / 1e3) / 1e9,
            'hbm_bound_ms': 403 * fpx / (peaks()[0] * 1e9) * 1e3 / world,
            'inputs': 'resident in HBM (uint8 BGR); results stay on the device'}
        del my_frames, batches
        # matcher
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'parrington.npz')) if os.path.exists(
            os.path.join(ROOT, 'tests', 'golden', 'parrington.npz')) else None
        NM = 65536
        if g is not None:
            pool = np.concatenate([g[f'desc_{i}'] for i in g['full_images'].tolist()])
            A_all, B_all = descriptor_sets('real', NM, NM, pool)
            dist_name = 'real-like: rows of the reference\'s parrington descriptors with +-2 jitter, seed 7'
        else:
            A_all, B_all = descriptor_sets('uniform', NM, NM)
            dist_name = 'uniform, seed 8'
        a_lo, a_hi = panorama.shard_range(NM, rank, world)
        A_loc = np.ascontiguousarray(A_all[a_lo:a_hi])
        msm = {}
        for name_e, top2 in (('best', 0), ('top2', 1)):
            ms = C.c_float()
            if world > 1:
                dist.barrier()
            _capi.check(ctx.lib.b200sift_bench_match(ctx.handle, _capi.ptr(A_loc), len(A_loc), _capi.ptr(B_all), NM, top2, 10,
                                                     C.byref(ms)))
            tm = torch.tensor([ms.value], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            msm[name_e] = float(tm[0])
        matcher_line = {'ms': msm, 'distribution': dist_name, 'nA': NM, 'nB': NM,
                        'split': f'A rows in contiguous blocks over {world} rank(s), B replicated'}
        fpipe.close()

    counts = np.asarray(out_e2e[1])
    desc_pairs = float(sum(int(counts[i]) * int(counts[i + 1]) for i in range(n - 1)))
    h2d = n * int(pinned_base[0].numel())                                   # whole job, all ranks
    d2h = int(counts.sum()) * (24 + 128) + (n - 1) * 16

    line = None
    if rank == 0:
        # ---- roofline of the dominant dense kernel: the separable blur at this workload's octave-0
        # layer shape (18 x 1024 x 768 float32 per launch), timed alone with CUDA events.
        peak, peak_src = peaks()
        lib = ctx.lib
        detail = {}
        sig = sift_impl.generate_gaussian_kernels(1.6, 3)
        t_sum = 0.0
        for name, s in [('base', 1.2489996)] + [(f'layer{l}', float(sig[l])) for l in range(1, 6)]:
            ms = C.c_float()
            _capi.check(lib.b200sift_bench_blur(ctx.handle, nb, 2 * h, 2 * w, s, 20, 1, C.byref(ms)))
            by = 8.0 * nb * (2 * h) * (2 * w)
            detail[name] = {'sigma': round(s, 4), 'ms': ms.value, 'GB/s': by / ms.value / 1e6}
            t_sum += ms.value
        by = 8.0 * nb * (2 * h) * (2 * w)
        ach = by / (t_sum / 6) / 1e6      # bytes per launch / average launch duration over the 6 blurs
        # the same bytes through a plain device copy (torch's copy kernel, the one MEASURED_PEAKS.json's
        # hbm_gbs was taken with at 1 Gi elements) under the same protocol: what a launch of this size
        # can reach at all (start-up, DRAM page opening and the drain are a fixed few microseconds)
        ca = torch.empty(nb * 2 * h * 2 * w, dtype=torch.float32, device=dev).normal_()
        cb = torch.empty_like(ca)
        for _ in range(3):
            cb.copy_(ca)
        t_copy = 0.0
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); cb.copy_(ca); e1.record()
            torch.cuda.synchronize()
            t_copy += e0.elapsed_time(e1) / 10
        del ca, cb
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of blur_ring_kernel at this shape: read from
        # the committed ncu summary of the same kernel (profiles/r1_final_ring_small_ncu.txt, unchanged this
        # round); the output is still dirty in L2 when the kernel ends, hence less than the algorithmic bytes
        traffic, traffic_src = ncu_traffic(os.path.join(ROOT, 'profiles', 'r1_final_ring_small_ncu.txt'))
        # the same kernel where the layer no longer fits one wave of CTAs (the 4096x3072-frame regime of
        # BASELINE.json configs[3]: 8 frames, octave-0 layers of 6144 x 8192): it is HBM-bound there
        large = {}
        for name, s in [('layer1', float(sig[1])), ('layer3', float(sig[3])), ('layer5', float(sig[5]))]:
            ms = C.c_float()
            _capi.check(lib.b200sift_bench_blur(ctx.handle, 8, 6144, 8192, s, 5, 1, C.byref(ms)))
            large[name] = {'sigma': round(s, 4), 'ms': ms.value, 'GB/s': 8.0 * 8 * 6144 * 8192 / ms.value / 1e6,
                           'frac': 8.0 * 8 * 6144 * 8192 / ms.value / 1e6 / peak}
        roof = {'bound': 'hbm', 'kernel': 'blur_ring_kernel<R> (octave-0 layer shape; average over the base blur '
                                          'and the 5 layer blurs of one octave: R = 5, 5, 6, 8, 10, 13)',
                'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                'traffic': (sum(traffic[r] for r in (5, 5, 6, 8, 10, 13)) / 6) if traffic else None,
                'traffic_source': traffic_src,
                'copy_same_bytes': {'ms': t_copy, 'GB/s': by / t_copy / 1e6, 'frac_of_peak': by / t_copy / 1e6 / peak,
                                    'what': 'torch copy_ of 18 x 1024 x 768 float32 (same 8 B/px), L2 flushed: the ceiling '
                                            'of a launch of this size'},
                'algorithmic_bytes_per_launch': 8 * nb * 2 * h * 2 * w, 'peak_source': peak_src,
                'timing': 'each launch alone between CUDA events on the launch stream, 256 MiB L2 flush before it',
                'per_sigma': detail, 'large_shape_8x6144x8192': large}
        # ---- the dominant kernel of the step is describe_kernel (28 % of the device time): bound by
        # instruction issue, not by HBM or the tensor pipe.  Its duration is measured here (CUDA events around
        # the launch, one step at a time, L2 flushed); the warp-instruction count of the same launch comes
        # from the committed ncu capture (profiles/r2_describe_ncu.txt: same inputs, deterministic).
        t_desc, n_desc = 0.0, 0
        for i in range(5):
            flush.fill_(i)
            torch.cuda.synchronize()
            sift_impl.detect_and_describe_batch(resident_base, ctx=ctx, download=False)
            ms = C.c_float()
            nk = C.c_int32()
            _capi.check(lib.b200sift_last_describe_ms(ctx.handle, C.byref(ms), C.byref(nk)))
            t_desc += ms.value / 5
            n_desc = nk.value
        winst = None
        pth = os.path.join(ROOT, 'profiles', 'r2_describe_ncu.txt')
        if os.path.exists(pth):
            import re
            mm = re.search(r'smsp__inst_executed\.sum\s+([0-9.]+) inst', open(pth).read())
            winst = float(mm.group(1)) if mm else None
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sm_count * 4 * (clocks['sm_max_mhz'] if clocks and clocks.get('sm_max_mhz') else 1965.0) * 1e6
        roof_desc = {'bound': 'issue', 'kernel': 'describe_kernel (4x4x8 descriptors, one warp per oriented keypoint)',
                     'ms': t_desc, 'keypoints': n_desc, 'keypoints_per_s': n_desc / (t_desc * 1e-3) if t_desc else None,
                     'warp_instructions': winst, 'warp_instructions_source': 'profiles/r2_describe_ncu.txt (ncu --set full)',
                     'achieved': winst / (t_desc * 1e-3) / 1e9 if (winst and t_desc) else None,
                     'peak': issue_peak / 1e9, 'unit': 'G warp-instructions/s',
                     'frac': winst / (t_desc * 1e-3) / issue_peak if (winst and t_desc) else None,
                     'peak_source': f'{sm_count} SMs x 4 schedulers x 1 warp instruction per cycle at the maximum SM clock'}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms_cpu, sample, threads, dps = cpu_run(imgs, 1, 1, budget_s=40.0)
            cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
                   'match_desc_pairs_per_s': dps,
                   'python_reference': 'not timed in this run: the reference is pure Python and does not travel to the GPU '
                                       'box (/root/reference is absent there); BASELINE.md holds its timings measured in '
                                       'the build container: 0.005-0.008 Mpix/s and 3.5-3.9e5 descriptor pairs/s on 8 cores'}
        line = {
            'metric': METRIC, 'value': mpix_step / (ms_dev / 1e3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': data,
            'config': {'workload': WORKLOAD,
                       'chain': f'{world} such set(s) chained in pano order ({n} images, {n - 1} pairs): per-GPU work is fixed',
                       'images': n, 'pairs': n - 1, 'keypoints': int(counts.sum()),
                       'sharding': f'contiguous blocks of {nb} images per rank over {world} rank(s); pair (i,i+1) on the '
                                   'owner of i; every rank sends its first image\'s descriptors to the previous rank (NCCL send / recv), '
                                   'shifts and counts are all-gathered (24 B per image)',
                       'mode': f'throughput: {DEPTH} steps in flight per GPU ({DEPTH} library contexts per rank; N = 1: '
                               'pipeline.PanoramaPipeline, N > 1: panorama.sharded_panorama_stream); device time over all '
                               'steps / steps.  single_step = one step at a time, per-step CUDA events',
                       'l2': 'inputs larger than L2: per-step pyramid working set ~0.47 GB per 18 images > 126 MB L2 '
                             '(throughput legs, no flush); the single_step leg writes a 256 MiB flush before every step'},
            'e2e': {'value': mpix_step / (ms_e2e / 1e3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'wall_ms_per_step': wall_e2e},
            'gpu_launches': launches, 'wall_ms_per_step': wall_dev, 'host_cpus': os.cpu_count(),
            'single_step': {'ms_per_step': ms_one, 'value': mpix_step / (ms_one / 1e3), 'unit': UNIT,
                            'gpu_launches': launches_one},
            'match_desc_pairs_per_s': desc_pairs / (ms_dev / 1e3), 'image_pairs_per_s': (n - 1) / (ms_dev / 1e3),
            'roofline': roof, 'roofline_describe': roof_desc, 'cpu_baseline': cpu, 'clocks': clocks,
        }
        if strong:
            line['strong_18_images'] = strong
        if frames_line:
            line['frames_64x4096x3072'] = frames_line
        if matcher_line:
            pk8, pk8_src = int8_peak(torch, dev)
            ops = 2.0 * 128 * matcher_line['nA'] * matcher_line['nB']
            tops = {k: ops / (v * 1e-3) / 1e12 for k, v in matcher_line['ms'].items()}
            line['roofline_matcher'] = {
                'bound': 'tensor', 'kernel': 'match_tc_kernel<false> (tcgen05.mma kind::i8, nearest-neighbour epilogue)',
                'achieved': tops['best'], 'peak': pk8 * world, 'unit': 'TOP/s', 'frac': tops['best'] / (pk8 * world),
                'peak_source': pk8_src + (f' x {world} GPUs' if world > 1 else ''),
                'top2_epilogue': {'achieved': tops['top2'], 'frac': tops['top2'] / (pk8 * world)},
                'algorithmic_ops_per_launch': ops, 'desc_pairs_per_s': matcher_line['nA'] * matcher_line['nB'] /
                (matcher_line['ms']['best'] * 1e-3), 'frac_of_measured_bf16_peak': tops['best'] / (1662.8 * world),
                'timing': 'kernel alone, mean of 10 launches between CUDA events on the launch stream (max over ranks); '
                          'operands (8 MB + 8 MB) stay in L2 between launches by design: B is re-read by every A tile',
                **{k: matcher_line[k] for k in ('distribution', 'nA', 'nB', 'split', 'ms')}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        for be in backends:
            be.close()                      # exchange buffers go before the communicator and the contexts
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
