#!/bin/bash
# last validation of the round: the whole GPU suite, a quick default-workload bench line (clock sampler, e2e ring), compute-sanitizer on the kernels added late
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu_r02.log 2>&1; tail -4 $O/pytest_gpu_r02.log
timeout 300 python bench.py --no-hbm-record --no-cpu-baseline > $O/bench_r02_quick.json 2>$O/bench_r02_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_quick.json').read().strip().splitlines()[-1]); print('quick', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['clocks'])
PY
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or batching_and_sharding" > $O/sanitizer_memcheck_tail_mirror_r02.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/sanitizer_memcheck_tail_mirror_r02.log
