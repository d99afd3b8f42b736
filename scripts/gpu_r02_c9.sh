#!/bin/bash
# compute-sanitizer racecheck + initcheck on the kernels added late in the round (refilling tail kernel, mirror stores, four-slot ring)
O=gpurun_out; mkdir -p $O
for tool in racecheck initcheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or batching_and_sharding or async_tonemap" > $O/sanitizer_${tool}_tail_mirror_r02.log 2>&1; echo "$tool rc=$?"; tail -4 $O/sanitizer_${tool}_tail_mirror_r02.log
done
