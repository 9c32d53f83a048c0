#!/bin/bash
# the 64-frame leg of bench.py: the current library, then libb200sift_variant.so copied over it
mkdir -p gpurun_out
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$1', 'value ms', d['ms_per_step'], 'frames ms', d['frames_64x4096x3072']['ms'], 'kp', d['frames_64x4096x3072']['keypoints'], 'matcher', d['roofline_matcher'].get('ms'))"
}
run current 2>&1 | tee gpurun_out/frames_leg_current.txt
run current2 2>&1 | tee -a gpurun_out/frames_leg_current.txt
cp vfx_image_stitching_b200/libb200sift_variant.so vfx_image_stitching_b200/libb200sift.so
run variant 2>&1 | tee gpurun_out/frames_leg_variant.txt
