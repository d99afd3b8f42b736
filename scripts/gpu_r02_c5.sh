#!/bin/bash
# validation of the final library: smoke, the whole GPU suite, the default bench line (what the driver runs)
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02.log 2>&1; tail -2 $O/smoke_r02.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02.log 2>&1; tail -8 $O/pytest_gpu_r02.log
timeout 900 python bench.py > $O/bench_r02_default_flags.json 2>$O/bench_r02_default_flags.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_default_flags.json').read().strip().splitlines()[-1]); print('full default', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])
PY
