#!/bin/bash
# A/B helper (bring-up): bench under different env toggles / library builds.
# usage: WL=sponza_1080p STEPS=64 bash scripts/gpu_ab.sh "label|ENV=1 ENV2=x|extra bench args" ...
WL=${WL:-sponza_1080p}
for spec in "$@"; do
  IFS='|' read -r label envs extra <<< "$spec"
  env $envs python bench.py --workload $WL --steps ${STEPS:-64} --warmup 16 --no-cpu-baseline --no-e2e $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']; r=d['roofline']
print('%-28s Mrays/s=%.0f ms=%.1f | serial: closest=%.1f shadow=%.1f shade=%.1f sort=%.1f nodes/ray=%.1f build=%.0fms' % ('$label',d['value'],d['ms_per_step']*d['steps'],k['trace_closest'],k['trace_shadow'],k['shade'],k.get('sort',0.0),r['per_ray']['nodes'],d['config']['bvh_build']['ms']))"
done
