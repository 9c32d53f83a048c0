#!/bin/bash
# ring blur: input-stage depth / segment-height sweep (build-time macros)
for F in "" "-DB200SIFT_RING_S_SMALL=4" "-DB200SIFT_RING_S_SMALL=4 -DB200SIFT_RING_S_10=4 -DB200SIFT_RING_S_13=3" "-DB200SIFT_RING_S_SMALL=2" "-DB200SIFT_RING_SEG=64" "-DB200SIFT_RING_SEG=256"; do
  touch vfx_image_stitching_b200/csrc/blur_ring.cuh
  B200SIFT_NVCC_FLAGS="$F" python -m vfx_image_stitching_b200.build > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  echo "== flags: $F"
  python tools/blur_sweep.py
done
