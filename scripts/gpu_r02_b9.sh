#!/bin/bash
# Round-2 run B9: per-bounce counters (no swap kernel) + finish fused into accumulate: full suite, bench, interactive probe; hints A/B on the HBM-resident scene
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02.log 2>&1; tail -2 $O/smoke_r02.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02f.log 2>&1; tail -8 $O/pytest_gpu_r02f.log
timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02j_default.json 2>$O/bench_r02j_default.err
timeout 600 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02j_10m.json 2>/dev/null
VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_nohints.so timeout 600 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02j_10m_nohints.json 2>/dev/null
timeout 600 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02j_10m_b.json 2>/dev/null
for f in $O/bench_r02j_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), d['gpu_launches'], {k:round(v,2) for k,v in d['kernel_ms'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
timeout 300 python scripts/gpu_e2e_probe.py > $O/e2e_probe_r02b.log 2>&1; tail -14 $O/e2e_probe_r02b.log
