#!/bin/bash
# Round-2 run B5: alpha classification at commit + streaming cache hints (A/B), full suite
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02c.log 2>&1; tail -12 $O/pytest_gpu_r02c.log
timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02f_default.json 2>$O/bench_r02f_default.err
VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_nohints.so timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02f_nohints.json 2>/dev/null
VKRT_ALPHA_CLASSIFY=0 timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02f_noclassify.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02f_default2.json 2>/dev/null
for f in $O/bench_r02f_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), {k:round(v,2) for k,v in d['kernel_ms'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
