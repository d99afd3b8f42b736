#!/bin/bash
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02.log 2>&1; tail -2 $O/smoke_r02.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02g.log 2>&1; tail -8 $O/pytest_gpu_r02g.log
timeout 900 python bench.py > $O/bench_r02k_default.json 2>$O/bench_r02k_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02k_default.json').read().strip().splitlines()[-1]); print('full default', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
