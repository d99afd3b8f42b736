#!/bin/bash
# A/B helper (bring-up): bench with env-selected library variants / toggles
run() { python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']
print('$1  Mrays/s=%.0f wall=%.1f closest=%.1f shadow=%.1f shade=%.1f' % (d['value'],d['wall_ms'],k['trace_closest'],k['trace_shadow'],k['shade']))"; }
run base
for n in A B C D E; do VKRT_LIB_PATH=$PWD/vk_raytrace_b200/libvkrt_var_$n.so run $n; done
