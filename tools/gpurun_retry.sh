#!/bin/bash
# tools/gpurun_retry.sh <log> <gpurun args...> -- retries a gpurun call while the pod answers "busy" (nothing is charged for those)
log=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 150; else break; fi
done
