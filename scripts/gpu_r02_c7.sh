#!/bin/bash
# 8 GPUs, the driver's launch line: N = 1 on the same box for the efficiency, N = 8 with the peer exchange (full line), N = 8 with the gather exchange
O=gpurun_out; mkdir -p $O
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-hbm-record --no-e2e > $O/scale_r02b_n1.json 2> $O/scale_r02b_n1.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $O/scale_r02b_n8_peer.json 2> $O/scale_r02b_n8_peer.err
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --exchange gather --no-c4 --no-verify-exchange > $O/scale_r02b_n8_gather.json 2> $O/scale_r02b_n8_gather.err
for f in $O/scale_r02b_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], 'value', round(d.get('value'),1), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), 'ms/step', round(d.get('ms_per_step'),4), 'exchange ms', round(d.get('reduce_ms',0),3), 'verified', d['config'].get('exchange_verified'), 'c4', (d.get('c4') or {}).get('value'), (d.get('c4') or {}).get('exchange_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -c 300 $O/scale_r02b_n8_peer.err
