#!/bin/bash
# Round-2 run A (one B200): smoke, GPU tests, convergence curve, default bench, A/B of the round-2 traversal changes
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02.log 2>&1; tail -2 $O/smoke_r02.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/pytest_gpu_r02a.log 2>&1; tail -15 $O/pytest_gpu_r02a.log
timeout 600 python scripts/gpu_converge.py > $O/converge_r02.log 2>&1; tail -3 $O/converge_r02.log
timeout 900 python bench.py > $O/bench_r02a_default.json 2> $O/bench_r02a_default.err; tail -c 600 $O/bench_r02a_default.err
VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_nobf.so timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02a_nobf.json 2>/dev/null
VKRT_DUAL_SLOT=0 timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02a_single_slot.json 2>/dev/null
for f in $O/bench_r02a_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), round(d.get('ms_per_step'),3), {k:round(v,2) for k,v in d['kernel_ms'].items()}, (d.get('roofline') or {}).get('frac'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
