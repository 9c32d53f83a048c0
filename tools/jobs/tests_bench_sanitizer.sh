#!/bin/bash
# Round-2 first GPU job: full -m gpu suite, baseline bench, sanitizer logs.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2_sanitizer_$tool.log
done
tail -5 gpurun_out/r2_pytest_gpu.log
head -c 600 gpurun_out/r2_bench0.json
for tool in memcheck racecheck synccheck; do tail -4 gpurun_out/r2_sanitizer_$tool.log; done
