#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod reports "busy" (nothing is charged for those)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpu_try.log 2>&1
  if grep -q "status=transient" /tmp/gpu_try.log; then sleep 45; else break; fi
done
cat /tmp/gpu_try.log
