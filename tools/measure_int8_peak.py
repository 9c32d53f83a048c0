#!/usr/bin/env python
"""Dense int8 tensor-core peak of this GPU, the denominator of the matcher's roofline fraction
(MEASURED_PEAKS.json has only a bf16 figure; the matcher issues tcgen05.mma kind::i8).

cuBLASLt IGEMM through torch._int_mm (int8 x int8 -> int32), 8192^3, like the driver's bf16 probe:
best of 10 single launches (burst) and back to back for 2 s (sustained), CUDA events.
Also repeats the bf16 probe so that both numbers come from the same box and the same minute.

    python tools/measure_int8_peak.py [--out profiles/r2_int8_peak.json]
"""
import argparse
import json
import time

import torch


def probe(fn, n):
    flops = 2.0 * n ** 3
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 2.0:
        for _ in range(20):
            fn()
        k += 20
    e1.record()
    torch.cuda.synchronize()
    return flops / (best * 1e-3) / 1e12, flops * k / (e0.elapsed_time(e1) * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--n', type=int, default=8192)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    n = a.n
    out = {'gpu': torch.cuda.get_device_name(0), 'n': n, 'torch': torch.__version__}
    A = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    B = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev).t()   # column-major B, as cuBLASLt IGEMM wants
    try:
        burst, sust = probe(lambda: torch._int_mm(A, B), n)
        out.update(int8_tops=burst, int8_tops_sustained=sust, how='torch._int_mm (cuBLASLt IGEMM s8 x s8 -> s32)')
    except Exception as e:   # noqa: BLE001
        out['int8_error'] = repr(e)[:300]
    a16 = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    b16 = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    burst, sust = probe(lambda: torch.matmul(a16, b16), n)
    out.update(bf16_tflops=burst, bf16_tflops_sustained=sust)
    try:
        a8 = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn)
        b8 = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn).t()
        one = torch.tensor(1.0, device=dev)
        burst, sust = probe(lambda: torch._scaled_mm(a8, b8, scale_a=one, scale_b=one, out_dtype=torch.bfloat16), n)
        out.update(fp8_tflops=burst, fp8_tflops_sustained=sust)
    except Exception as e:   # noqa: BLE001
        out['fp8_error'] = repr(e)[:300]
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
