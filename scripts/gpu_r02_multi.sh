#!/bin/bash
# Round-2 multi-GPU run (N GPUs of one box): the driver's launch line for N = 1 .. $1
O=gpurun_out; mkdir -p $O
N=${1:-2}
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-hbm-record > $O/scale_r02_n1.json 2> $O/scale_r02_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 > $O/scale_r02_n$n.json 2> $O/scale_r02_n$n.err
    tail -c 400 $O/scale_r02_n$n.err
  fi
done
for f in $O/scale_r02_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], 'value', round(d.get('value'),1), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), 'ms/step', round(d.get('ms_per_step'),3), 'exchange ms', round(d.get('reduce_ms',0),3), 'bytes', d.get('exchange_bytes_per_rank'), 'c4', (d.get('c4') or {}).get('value'), (d.get('c4') or {}).get('exchange_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
