#!/bin/bash
# Round-2 run B7: cheap traversal-parameter A/Bs (each bit-identical by construction)
O=gpurun_out; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02h_$name.json 2>/dev/null; }
run base A=1
run stack16 VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_stack16.so
run refetch28 VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_refetch28.so
run refetch16 VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_refetch16.so
run stack16r28 VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_stack16r28.so
run leaf2 VKRT_SAH_LEAF=2
run leaf3 VKRT_SAH_LEAF=3
run leaf8 VKRT_SAH_LEAF=8
run base2 A=1
for f in $O/bench_r02h_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), {k:round(v,2) for k,v in d['kernel_ms'].items()}, {k:round(v,2) for k,v in d['roofline']['per_ray'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
timeout 900 python bench.py > $O/bench_r02h_full_default.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02h_full_default.json').read().strip().splitlines()[-1]); print('full default', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'])
PY
