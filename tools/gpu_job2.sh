#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "descriptors or end_to_end or full_set or out_pair or smoke or batch_equals" 2>&1 | tail -15 > gpurun_out/r2_desc_pytest.log
tail -6 gpurun_out/r2_desc_pytest.log
grep -h "descriptors\|e2e\|full set\|   parr\|   grail" gpurun_out/parity_report.txt | tail -30
B200SIFT_TRACE=1 python - <<'PY' 2>&1 | tail -12
import numpy as np, sys
sys.path.insert(0,'.')
from vfx_image_stitching_b200 import sift_impl as si
g=np.load('tests/golden/parrington.npz')['gray']
imgs=[np.ascontiguousarray(np.repeat(im[:,:,None],3,axis=2)) for im in g]
for _ in range(3): si.detect_and_describe_batch(imgs, download=False)
PY
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['single_step']['ms_per_step'])
PY
