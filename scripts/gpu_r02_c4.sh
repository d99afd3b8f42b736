#!/bin/bash
# fused tail with per-lane refill (in-tree) and its 4-blocks-per-SM build (build_alt): bit-exactness first, then interactive e2e and a rank's share of 8
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "tail or interactive or async" > $O/pytest_tail.log 2>&1; tail -3 $O/pytest_tail.log
for lib in refill tail4; do
  if [ $lib = tail4 ]; then export VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_tail4.so; else unset VKRT_LIB_PATH; fi
  timeout 600 python bench.py --no-hbm-record --no-cpu-baseline > $O/tailrefill_$lib.json 2>$O/tailrefill_$lib.err
  python - $O/tailrefill_$lib.json $lib <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as e: print(sys.argv[2], 'failed', e)
PY
  SHARD_PROBE_SHARDS=8,4 SHARD_PROBE_TAILS=0,-1,2 SHARD_PROBE_FIFS=20 timeout 300 python scripts/gpu_shard_probe.py > $O/shard_probe_$lib.log 2>&1; cat $O/shard_probe_$lib.log
done
