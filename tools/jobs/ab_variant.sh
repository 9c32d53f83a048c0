#!/bin/bash
# A/B of a build-time variant: the default library, then vfx_image_stitching_b200/libb200sift_variant.so
# (built here with B200SIFT_NVCC_FLAGS=-D..., copied over the default on the box).  GREP = timeline rows to show.
mkdir -p gpurun_out
G=${GREP:-"tail|refine|orient|describe"}
run() {
  echo "=== $1"
  python -m pytest tests -q -m gpu -x ${2} 2>&1 | tail -2
  for i in 1 2; do
  python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'single ms', d['single_step']['ms_per_step'], 'describe ms', d['roofline_describe']['ms'])"
  done
  B200SIFT_TIMELINE=1 python tools/profile_step.py 2>&1 | grep -E "$G" | tail -8
}
run default "-k full_set" 2>&1 | tee gpurun_out/ab_default.txt
cp vfx_image_stitching_b200/libb200sift_variant.so vfx_image_stitching_b200/libb200sift.so
run variant "" 2>&1 | tee gpurun_out/ab_variant.txt
