#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "transient/busy" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.out 2>&1
  if grep -q "status=transient\|status=refused" /tmp/gpurun_last.out; then sleep 45; continue; fi
  break
done
cat /tmp/gpurun_last.out
