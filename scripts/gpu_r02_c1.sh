#!/bin/bash
# A/B of the 64-byte path record (in-tree library) against the round's previous layout (build_alt/libvkrt_b200_soa.so), then the parity suite
O=gpurun_out; mkdir -p $O
ab() {  # $1 = tag, rest = env
  for lib in new soa; do
    if [ $lib = soa ]; then export VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_soa.so; else unset VKRT_LIB_PATH; fi
    env "${@:2}" timeout 600 python bench.py --no-hbm-record --no-cpu-baseline > $O/ab_$1_$lib.json 2>$O/ab_$1_$lib.err
    python - $O/ab_$1_$lib.json $1 $lib <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], sys.argv[3], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['ms_per_step'])
except Exception as e: print(sys.argv[2], sys.argv[3], 'failed', e)
PY
  done
  unset VKRT_LIB_PATH
}
ab default A=1
ab default2 A=1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/pytest_gpu_r02h.log 2>&1; tail -8 $O/pytest_gpu_r02h.log
