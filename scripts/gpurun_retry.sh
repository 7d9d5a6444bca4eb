#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout_s> <command string>  -- retries while the pod answers busy (rc 3)
log=$1; to=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "[retry] done rc=$rc attempts=$attempt" >> $log
