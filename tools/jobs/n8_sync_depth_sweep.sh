#!/bin/bash
# N = 8: host-wait mode x depth
for CFG in "0 4" "1 4" "1 3" "0 3"; do
  set -- $CFG
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 20 --warmup 3 --sync-mode $1 --depth $2 --no-extra 2>/tmp/b.err > /tmp/b.json || tail -5 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('sync=$1 depth=$2', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'single', round(d['single_step']['ms_per_step'],4), 'strong', round(d['strong_18_images']['ms_per_step'],4))"
done
