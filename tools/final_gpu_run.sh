#!/bin/bash
# Round-end measurement pass on one B200 (run through gpurun); everything lands in gpurun_out/.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r1_pytest_gpu.log 2>&1; tail -3 $O/r1_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r1_final_bench.json 2> $O/r1_final_bench.err; cut -c1-300 $O/r1_final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r1_final_reference_arm.json 2>> $O/r1_final_bench.err; cut -c1-300 $O/r1_final_reference_arm.json
python bench_configs.py grail out frames --frames 8 --steps 5 --check --out $O/r1_final_configs.json > /dev/null 2> $O/configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:describe_kernel --launch-skip 3 -c 1 -o $O/r1_final_describe python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_desc.log 2>&1
# one launch per radius (blur_probe: 3 warm-ups + 1 timed launch per sigma -> launch 4i+3); the .ncu-rep files
# must stay small: gpurun copies at most 64 MiB back
for i in 0 4; do
  ncu --set full --clock-control none --import-source on -k regex:blur_ring --launch-skip $((4*i+3)) -c 1 -o $O/r1_final_ring_small_s$i python tools/blur_probe.py 18 1024 768 1 > $O/ncu_ring_small.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:blur_ring --launch-skip $((4*i+3)) -c 1 -o $O/r1_final_ring_large_s$i python tools/blur_probe.py 8 6144 8192 1 > $O/ncu_ring_large.log 2>&1
done
ncu --set full --clock-control none -k regex:blur_ring -c 20 --csv --page raw --log-file $O/r1_final_ring_small_raw.csv python tools/blur_probe.py 18 1024 768 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:blur_ring -c 20 --csv --page raw --log-file $O/r1_final_ring_large_raw.csv python tools/blur_probe.py 8 6144 8192 1 > /dev/null 2>&1
du -sh $O; ls -la $O | tail -20
