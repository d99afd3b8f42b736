#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "transient / busy" (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -70
  if echo "$out" | grep -q "status=transient\|exit code 3\|status=busy"; then echo "[retry $i] waiting"; sleep 150; continue; fi
  break
done
