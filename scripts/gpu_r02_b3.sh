#!/bin/bash
# Round-2 run B3: builders again (group-aware leaves, warm build time), ncu captures with the right skip (fused tail off: 16 trace launches per batch)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --tb=short -x -k "trace_rays_bit_exact or full_size" > $O/pytest_gpu_r02b3.log 2>&1; tail -4 $O/pytest_gpu_r02b3.log
for mode in 2 4 1; do
  timeout 600 python bench.py --build-mode $mode --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02d_mode$mode.json 2>$O/bench_r02d_mode$mode.err
done
VKRT_SAH_STOP=2 timeout 600 python bench.py --build-mode 4 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02d_mode4_stop2.json 2>/dev/null
VKRT_SAH_STOP=8 VKRT_SAH_LEAF=8 timeout 600 python bench.py --build-mode 4 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02d_mode4_stop8.json 2>/dev/null
for f in $O/bench_r02d_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), d['config']['bvh_build']['ms'], d['config']['bvh_build']['ms_first_build_in_process'], {k:round(v,2) for k,v in d['kernel_ms'].items()}, {k:round(v,2) for k,v in d['roofline']['per_ray'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
export VKRT_TAIL_FROM=0
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 32 -c 4 -o $O/prof_trace_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceKernelWide -s 32 -c 2 -o $O/prof_trace10m_r02 -f python bench.py --workload instanced_10m --steps 8 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace10m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 16 -c 2 -o $O/prof_shade_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_shade.log 2>&1
ls -la $O/*r02*.ncu-rep
