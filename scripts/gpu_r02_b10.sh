#!/bin/bash
# Round-2 run B10: fused tail in batched mode; final ncu launch list + captures of the final kernels
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gpu_tail_probe.py > $O/tail_probe_r02.log 2>&1; cat $O/tail_probe_r02.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_r02.csv python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_list.log 2>&1
export VKRT_TAIL_FROM=0
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 32 -c 4 -o $O/prof_trace_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceKernelWide -s 32 -c 2 -o $O/prof_trace10m_r02 -f python bench.py --workload instanced_10m --steps 8 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_trace10m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 16 -c 2 -o $O/prof_shade_r02 -f python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-hbm-record > $O/ncu_shade.log 2>&1
ls -la $O/*r02*.ncu-rep
