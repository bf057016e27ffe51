#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout-seconds> <command...>   (retries while the pod answers "transient")
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy" $log; then sleep 90; else break; fi
done
