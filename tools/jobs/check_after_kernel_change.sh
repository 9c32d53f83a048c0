#!/bin/bash
# the GPU parity suite + a short bench + the per-phase timeline of one step (after a kernel change)
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/check_pytest.log
python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>gpurun_out/check_bench.err | tee gpurun_out/check_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'single ms', d['single_step']['ms_per_step'], 'describe ms', d['roofline_describe']['ms'])"
B200SIFT_TIMELINE=1 python tools/profile_step.py 2>&1 | grep -E "orient|refine|describe|counters|gather" | tail -6 | tee gpurun_out/check_timeline.txt
