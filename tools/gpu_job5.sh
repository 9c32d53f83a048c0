#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "matcher or match or ratio or pairs or shift or smoke or full_set" 2>&1 | tail -5
timeout 300 python tools/measure_int8_peak.py --out gpurun_out/r2_int8_peak.json
timeout 600 python bench_matcher.py --out gpurun_out/r2_matcher_sweep.json 2>&1 | tail -30
