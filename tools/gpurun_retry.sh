#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod reports "busy"
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpu_try.log 2>&1
  if grep -q "status=transient" /tmp/gpu_try.log; then sleep 90; else break; fi
done
cat /tmp/gpu_try.log
