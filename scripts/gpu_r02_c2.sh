#!/bin/bash
# wavefront slots: 2 (round default so far) against 3 and 4, batched headline + interactive e2e; then the parity suite with 4 slots
O=gpurun_out; mkdir -p $O
for n in 2 3 4 2 4; do
  VKRT_SLOTS=$n timeout 600 python bench.py --no-hbm-record --no-cpu-baseline > $O/slots_$n.json 2>$O/slots_$n.err
  python - $O/slots_$n.json $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('slots', sys.argv[2], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['ms_per_step'], d['e2e']['ms_per_step'])
except Exception as e: print(sys.argv[2], 'failed', e)
PY
done
VKRT_SLOTS=4 timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/pytest_gpu_r02i.log 2>&1; tail -8 $O/pytest_gpu_r02i.log
