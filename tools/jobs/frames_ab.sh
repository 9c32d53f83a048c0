#!/bin/bash
# per-phase device times (B200SIFT_TRACE) of detect+describe on 4 synthetic 4096x3072 frames: the current
# library, then vfx_image_stitching_b200/libb200sift_variant.so copied over it
mkdir -p gpurun_out
run() {
  echo "=== $1"
  B200SIFT_TRACE=1 python tools/profile_step.py --frames 4 2>&1 | grep -E "trace|ok" | tail -9
}
run current 2>&1 | tee gpurun_out/frames_ab_current.txt
cp vfx_image_stitching_b200/libb200sift_variant.so vfx_image_stitching_b200/libb200sift.so
run variant 2>&1 | tee gpurun_out/frames_ab_variant.txt
