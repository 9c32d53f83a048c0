#!/bin/bash
N=${1:-2}
for D in 3 4 5; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --depth $D --no-extra --no-cpu-baseline 2>/tmp/b.err > /tmp/b.json || tail -5 /tmp/b.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 --depth $D --no-extra 2>/tmp/b.err > /tmp/b.json || tail -5 /tmp/b.err
  fi
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('N=$N depth=$D', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'single', round(d['single_step']['ms_per_step'],4))"
done
