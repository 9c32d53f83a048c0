#!/bin/bash
# Round-end measurement pass on one B200 (run through gpurun); everything lands in gpurun_out/.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/r2_pytest_gpu.log 2>&1; tail -3 $O/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; cut -c1-200 $O/r2_final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_reference_arm.json 2>> $O/r2_final_bench.err; cut -c1-200 $O/r2_final_reference_arm.json
python bench_configs.py grail out --steps 5 --out $O/r2_final_configs.json > /dev/null 2> $O/configs.err
python bench_matcher.py --out $O/r2_final_matcher_sweep.json > /dev/null 2>&1
# ncu launch list of the bench command (headline legs only: the frames / matcher legs add thousands of launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_bench.log 2>&1
python profiles/summarize_launches.py $O/r2_final_launches.csv > $O/r2_final_summary_launches.txt 2>&1; head -30 $O/r2_final_summary_launches.txt
# one step, every kernel, ncu --set full (raw page only; the report itself stays on the box)
ncu --set full --clock-control none --profile-from-start off --csv --page raw --log-file $O/r2_step_full_raw.csv python tools/profile_step.py > /dev/null 2>&1
python profiles/summarize_ncu_raw.py $O/r2_step_full_raw.csv > $O/r2_step_full_ncu.txt 2>&1
# the matcher at 64k x 64k
ncu --set full --clock-control none -k regex:match_tc_kernel --launch-skip 3 -c 1 --csv --page raw --log-file $O/r2_matcher_raw.csv python -c "
import ctypes as C, sys
sys.path.insert(0,'.')
from vfx_image_stitching_b200 import _capi
ctx=_capi.default_context(0); ms=C.c_float()
_capi.check(ctx.lib.b200sift_bench_match(ctx.handle, None, 65536, None, 65536, 0, 3, C.byref(ms)))
" > /dev/null 2>&1
python profiles/summarize_ncu_raw.py $O/r2_matcher_raw.csv > $O/r2_matcher_ncu.txt 2>&1; cat $O/r2_matcher_ncu.txt
du -sh $O
