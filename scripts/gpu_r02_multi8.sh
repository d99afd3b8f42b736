#!/bin/bash
# Round-2 8-GPU run: tile size of the shard interleave (load balance) at N = 8
O=gpurun_out; mkdir -p $O
for tile in 64 32 16; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+tile)) bench.py --gpus 8 --steps 20 --warmup 3 --tile $tile --no-c4 --no-e2e > $O/scale_r02_n8_tile$tile.json 2> $O/scale_r02_n8_tile$tile.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 64 --warmup 4 --no-e2e > $O/scale_r02_n8_steps64.json 2> $O/scale_r02_n8_steps64.err
for f in $O/scale_r02_n8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], 'value', round(d.get('value'),1), 'ms/step', round(d.get('ms_per_step'),4), 'exchange ms', round(d.get('reduce_ms',0),3), d['kernel_ms'], (d.get('c4') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
