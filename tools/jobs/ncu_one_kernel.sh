#!/bin/bash
# ncu --set full of one kernel of the headline step; KERNEL=regex NAME=tag
mkdir -p gpurun_out
K=${KERNEL:-describe_kernel}
N=${NAME:-r2_describe}
python tools/profile_step.py > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c ${COUNT:-1} \
    -o gpurun_out/$N -f python tools/profile_step.py > gpurun_out/${N}_ncu.log 2>&1
tail -3 gpurun_out/${N}_ncu.log
ncu -i gpurun_out/$N.ncu-rep --page raw --csv > gpurun_out/${N}_raw.csv 2>/dev/null
python profiles/summarize_ncu_raw.py gpurun_out/${N}_raw.csv
