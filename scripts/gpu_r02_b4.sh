#!/bin/bash
# Round-2 run B4: sync-free GPU SAH builder with pooled temporaries
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "trace_rays_bit_exact or full_size or timed_configuration or degenerate" > $O/pytest_gpu_r02b4.log 2>&1; tail -4 $O/pytest_gpu_r02b4.log
for mode in 4 2 1 3; do
  timeout 600 python bench.py --build-mode $mode --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02e_mode$mode.json 2>$O/bench_r02e_mode$mode.err
done
timeout 900 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-e2e > $O/bench_r02e_instanced_10m.json 2>/dev/null
timeout 900 python bench.py --workload helmet_1080p --no-cpu-baseline --no-hbm-record > $O/bench_r02e_helmet_1080p.json 2>/dev/null
for f in $O/bench_r02e_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round(d['config']['bvh_build']['ms'],2), round(d['config']['bvh_build']['ms_first_build_in_process'],1), {k:round(v,2) for k,v in d['kernel_ms'].items()}, {k:round(v,2) for k,v in d['roofline']['per_ray'].items()}, d['roofline'].get('frac'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
