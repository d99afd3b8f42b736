#!/bin/bash
# Round-2 run B8: device-side ray sort, scene-file path, shard scheduling probe, full default bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "batching or scene_file or async_tonemap or cornell_matches or sponza_with_alpha" > $O/pytest_gpu_r02e.log 2>&1; tail -5 $O/pytest_gpu_r02e.log
timeout 600 python bench.py --no-cpu-baseline --no-hbm-record > $O/bench_r02i_default.json 2>$O/bench_r02i_default.err
timeout 600 python bench.py --no-cpu-baseline --no-hbm-record --sort-rays > $O/bench_r02i_sort.json 2>$O/bench_r02i_sort.err
timeout 600 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-hbm-record --no-e2e > $O/bench_r02i_10m.json 2>/dev/null
timeout 600 python bench.py --workload instanced_10m --steps 16 --no-cpu-baseline --no-hbm-record --no-e2e --sort-rays > $O/bench_r02i_10m_sort.json 2>/dev/null
for f in $O/bench_r02i_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d.get('value'),1), round((d.get('e2e') or {}).get('value') or 0,1), {k:round(v,2) for k,v in d['kernel_ms'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
timeout 300 python scripts/gpu_shard_probe.py > $O/shard_probe_dual_r02.log 2>&1; cat $O/shard_probe_dual_r02.log
VKRT_DUAL_SLOT=0 timeout 300 python scripts/gpu_shard_probe.py > $O/shard_probe_single_r02.log 2>&1; cat $O/shard_probe_single_r02.log
timeout 900 python bench.py > $O/bench_r02i_full_default.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02i_full_default.json').read().strip().splitlines()[-1]); print('full default', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['traffic'])
PY
