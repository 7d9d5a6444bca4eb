#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <ngpus> <logfile> <timeout_s> <command string>
n=$1; log=$2; to=$3; shift 3
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "[retry] done rc=$rc attempts=$attempt" >> $log
