#!/bin/bash
mkdir -p gpurun_out
B200SIFT_TIMELINE=1 python tools/profile_step.py 2>&1 | grep timeline | tail -34 > gpurun_out/r2_timeline.txt
cat gpurun_out/r2_timeline.txt | grep -i "extrema\|refine\|orient\|tail\|describe\|layer 3\|layer 5"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_step_launches.csv python tools/profile_step.py > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_step_launches.csv 1 | head -40
