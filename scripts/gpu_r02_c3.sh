#!/bin/bash
# interactive frames with four slots: where should the fused tail start?  and a rank's share of an 8-GPU run with 2 / 3 / 4 slots
O=gpurun_out; mkdir -p $O
for t in -1 0 4 5; do
  if [ $t = -1 ]; then unset VKRT_TAIL_FROM; else export VKRT_TAIL_FROM=$t; fi
  timeout 600 python bench.py --no-hbm-record --no-cpu-baseline > $O/tail_$t.json 2>$O/tail_$t.err
  python - $O/tail_$t.json $t <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('tail_from', sys.argv[2], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as e: print(sys.argv[2], 'failed', e)
PY
done
unset VKRT_TAIL_FROM
for n in 2 3 4; do
  SHARD_PROBE_SHARDS=8 VKRT_SLOTS=$n timeout 300 python scripts/gpu_shard_probe.py > $O/shard_probe_slots$n.log 2>&1; cat $O/shard_probe_slots$n.log
done
