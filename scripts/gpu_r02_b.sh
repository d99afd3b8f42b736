#!/bin/bash
# Round-2 run B2 (one B200): full GPU suite, builders, interactive probe with / without the fused tail, ncu launch list + full captures
# (headline and HBM-resident workload), compute-sanitizer
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02b.log 2>&1; tail -12 $O/pytest_gpu_r02b.log
for mode in 2 4; do
  timeout 600 python bench.py --build-mode $mode --no-cpu-baseline --no-hbm-record > $O/bench_r02c_mode$mode.json 2>$O/bench_r02c_mode$mode.err
done
VKRT_SAH_LEAF=8 timeout 600 python bench.py --build-mode 4 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02c_mode4_leaf8.json 2>/dev/null
timeout 900 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --build-mode 4 > $O/bench_r02c_instanced_10m_mode4.json 2>/dev/null
for f in $O/bench_r02c_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), round(d['config']['bvh_build']['ms'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()}, {k:round(v,2) for k,v in d['roofline']['per_ray'].items()}, d['gpu_launches'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
timeout 300 python scripts/gpu_e2e_probe.py > $O/e2e_probe_tail_r02.log 2>&1; tail -14 $O/e2e_probe_tail_r02.log
VKRT_TAIL_FROM=0 timeout 300 python scripts/gpu_e2e_probe.py > $O/e2e_probe_notail_r02.log 2>&1; tail -14 $O/e2e_probe_notail_r02.log
# ncu: never a bench value.  Launch list of the default run, then full captures of the trace kernels of the first timed batch.
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_r02.csv python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 32 -c 4 -o $O/prof_trace_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceKernelWide -s 32 -c 2 -o $O/prof_trace10m_r02 -f python bench.py --workload instanced_10m --steps 8 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace10m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 16 -c 2 -o $O/prof_shade_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_shade.log 2>&1
ls -la $O/*r02*.ncu-rep
# compute-sanitizer (persistent kernels on two streams x two wavefront slots share device state): small configurations
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke_r02.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $O/sanitizer_memcheck_smoke_r02.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck_smoke_r02.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 $O/sanitizer_racecheck_smoke_r02.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sponza_with_alpha and 0-0 or async_tonemap or reference_branches and blend" > $O/sanitizer_memcheck_alpha_r02.log 2>&1; echo "memcheck alpha rc=$?"; tail -3 $O/sanitizer_memcheck_alpha_r02.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sponza_with_alpha and 0-0" > $O/sanitizer_racecheck_alpha_r02.log 2>&1; echo "racecheck alpha rc=$?"; tail -3 $O/sanitizer_racecheck_alpha_r02.log
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_initcheck_smoke_r02.log 2>&1; echo "initcheck smoke rc=$?"; tail -3 $O/sanitizer_initcheck_smoke_r02.log
