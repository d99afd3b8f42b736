#!/bin/bash
# Round-2 run B1: blend diagnosis, GPU SAH builder, leaf-stream A/B, interactive-mode probe
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "reference_branches or blend_probe or trace_rays_bit_exact or forced_tree or async_tonemap or timed_configuration or converged" > $O/pytest_gpu_r02b1.log 2>&1; tail -30 $O/pytest_gpu_r02b1.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "full_size" > $O/pytest_gpu_r02b1_full.log 2>&1; tail -8 $O/pytest_gpu_r02b1_full.log
for mode in 2 4; do
  timeout 600 python bench.py --build-mode $mode --no-cpu-baseline --no-hbm-record > $O/bench_r02b_mode$mode.json 2>$O/bench_r02b_mode$mode.err
done
VKRT_LIB_PATH=$PWD/build_alt/libvkrt_b200_nols.so timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02b_nols.json 2>/dev/null
VKRT_SAH_STOP=4 timeout 600 python bench.py --build-mode 4 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02b_mode4_stop4.json 2>/dev/null
VKRT_SAH_STOP=16 timeout 600 python bench.py --build-mode 4 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02b_mode4_stop16.json 2>/dev/null
for f in $O/bench_r02b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), round(d['config']['bvh_build']['ms'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()}, {k:round(v,2) for k,v in d['roofline']['per_ray'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
timeout 600 python scripts/gpu_e2e_probe.py > $O/e2e_probe_r02.log 2>&1; cat $O/e2e_probe_r02.log | tail -30
