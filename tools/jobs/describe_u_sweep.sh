#!/bin/bash
# describe kernel: U = 2 / 3 / 4 chunk slots per lane (build-time), per-phase trace + bench
mkdir -p gpurun_out
for U in 2 3 4; do
  touch vfx_image_stitching_b200/csrc/describe.cuh
  B200SIFT_NVCC_FLAGS="-DB200SIFT_DESC_U=$U" python -m vfx_image_stitching_b200.build > /dev/null 2>&1 || { echo build failed U=$U; continue; }
  echo "== U=$U"
  python -m pytest tests -m gpu -q -k "descriptors_match or full_set" 2>&1 | tail -1
  B200SIFT_TRACE=1 python - <<'PY' 2>&1 | grep describe | tail -2
import numpy as np, sys
sys.path.insert(0,'.')
from vfx_image_stitching_b200 import sift_impl as si
g=np.load('tests/golden/parrington.npz')['gray']
imgs=[np.ascontiguousarray(np.repeat(im[:,:,None],3,axis=2)) for im in g]
for _ in range(4): si.detect_and_describe_batch(imgs, download=False)
PY
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['single_step']['ms_per_step'])"
done
