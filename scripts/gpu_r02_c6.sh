#!/bin/bash
# 2 GPUs: the mirror test, then the driver's launch line with the peer exchange and with the gather exchange
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q --tb=short -x -k "mirror or async_tonemap" > $O/pytest_mirror.log 2>&1; tail -5 $O/pytest_mirror.log
for mode in peer gather; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $mode > $O/scale_r02_n2_$mode.json 2> $O/scale_r02_n2_$mode.err
  tail -c 600 $O/scale_r02_n2_$mode.err
  python - $O/scale_r02_n2_$mode.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], 'value', round(d.get('value'),1), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), 'ms/step', round(d.get('ms_per_step'),3), 'exchange ms', round(d.get('reduce_ms',0),3), 'verified', d['config'].get('exchange_verified'), d['config']['sharding'][:90], 'c4', (d.get('c4') or {}).get('value'), (d.get('c4') or {}).get('exchange_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
