#!/bin/bash
# Round-1 closing run (one B200): smoke, tests, bench records -> gpurun_out/  (ncu captures: scripts/gpu_final_r01.sh)
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r01.log 2>&1; tail -3 $O/pytest_gpu_r01.log
python bench.py > $O/bench_r01_default_flags.json 2>/dev/null
python bench.py --steps 64 --warmup 16 > $O/bench_r01_sponza_1080p.json 2> $O/bench_r01_sponza_1080p.err
python bench.py --workload instanced_10m --steps 32 --warmup 16 --no-cpu-baseline > $O/bench_r01_instanced_10m.json 2>/dev/null
python bench.py --workload helmet_1080p --steps 64 --warmup 16 --no-cpu-baseline > $O/bench_r01_helmet_1080p.json 2>/dev/null
python bench.py --workload sponza_4k --steps 16 --warmup 4 --no-cpu-baseline > $O/bench_r01_sponza_4k.json 2>/dev/null
for f in $O/bench_r01_default_flags.json $O/bench_r01_sponza_1080p.json $O/bench_r01_instanced_10m.json $O/bench_r01_helmet_1080p.json $O/bench_r01_sponza_4k.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value',0),1), d.get('ms_per_step'), d.get('clocks'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
