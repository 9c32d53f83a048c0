#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['single_step']['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['copy_same_bytes'], d['roofline']['traffic'])
print('frames', d.get('frames_64x4096x3072'))
print('matcher', d.get('roofline_matcher'))
PY
