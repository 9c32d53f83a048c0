#!/bin/bash
# multi-GPU job: N = $1 ; tests (N >= 2) + bench via torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$2" != "benchonly" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
fi
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2_nccl_n$N.%h.%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
grep -h "NVLS\|comm .* rank .* nranks\|Connected all" gpurun_out/r2_nccl_n$N.*.log | head -4
rm -f gpurun_out/r2_nccl_n$N.*.log
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'single', d['single_step']['ms_per_step'])
print('strong', d.get('strong_18_images'))
f=d.get('frames_64x4096x3072'); print('frames', f and {k:f[k] for k in ('ms','value','frames_per_rank')})
m=d.get('roofline_matcher'); print('matcher', m and {k:m[k] for k in ('ms','achieved','frac')})
PY
