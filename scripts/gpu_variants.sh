#!/bin/bash
# quick A/B of traversal-kernel variants and BVH build modes (bring-up helper)
for bm in 1; do for v in 8 200; do
  VKRT_TRACE_VARIANT=$v python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --build-mode $bm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms']; r=d['roofline']
print('build_mode=$bm variant=$v  Mrays/s=%.0f closest=%.1fms shadow=%.1fms shade=%.1fms nodes/ray=%.1f tris/ray=%.1f' % (d['value'],k['trace_closest'],k['trace_shadow'],k['shade'],r['per_ray']['nodes'],r['per_ray']['tris']))"
done; done
