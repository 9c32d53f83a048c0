#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B200SIFT_TRACE=1 python - <<'PY' 2>&1 | grep "describe\|orient" | tail -4
import numpy as np, sys
sys.path.insert(0,'.')
from vfx_image_stitching_b200 import sift_impl as si
g=np.load('tests/golden/parrington.npz')['gray']
imgs=[np.ascontiguousarray(np.repeat(im[:,:,None],3,axis=2)) for im in g]
for _ in range(4): si.detect_and_describe_batch(imgs, download=False)
PY
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['single_step']['ms_per_step'])"
NAME=r2_describe_b bash tools/gpu_ncu_desc.sh | tail -20
