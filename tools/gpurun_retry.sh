#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> '<command>'   -- retries while the pod answers "transient / no slot" (nothing charged)
log=$1; to=$2; cmd=$3
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$cmd" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "gpurun_retry: finished rc=$rc after $attempt attempt(s)" >> $log
