#!/bin/bash
# Round-1 evidence run (one B200): tests, bench records, ncu launch list and full captures -> gpurun_out/
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r01.log 2>&1; tail -3 $O/pytest_gpu_r01.log
python bench.py --steps 64 --warmup 16 > $O/bench_r01_sponza_1080p.json 2> $O/bench_r01_sponza_1080p.err
python bench.py --steps 64 --warmup 16 --build-mode 1 --no-cpu-baseline --no-e2e > $O/bench_r01_sponza_1080p_lbvh.json 2>/dev/null
python bench.py --steps 64 --warmup 16 --build-mode 3 --no-cpu-baseline --no-e2e > $O/bench_r01_sponza_1080p_ploc.json 2>/dev/null
python bench.py --workload instanced_10m --steps 32 --warmup 16 --no-cpu-baseline > $O/bench_r01_instanced_10m.json 2>/dev/null
python bench.py --workload helmet_1080p --steps 64 --warmup 16 --no-cpu-baseline > $O/bench_r01_helmet_1080p.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r01_reference_arm.json 2>/dev/null
for f in $O/bench_r01_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], d.get('value'), (d.get('e2e') or {}).get('value'), d.get('ms_per_step'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r01.csv python bench.py --steps 32 --warmup 32 --no-cpu-baseline --no-e2e > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 16 -c 4 -o $O/prof_trace_r01b -f python bench.py --steps 32 --warmup 32 --no-cpu-baseline --no-e2e > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 8 -c 2 -o $O/prof_shade_r01b -f python bench.py --steps 32 --warmup 32 --no-cpu-baseline --no-e2e > $O/ncu_shade.log 2>&1
ls -la $O/*.ncu-rep | tail -3
