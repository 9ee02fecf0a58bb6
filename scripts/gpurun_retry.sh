#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <script> [logfile]  -- retries while the pod answers "busy" (exit code 3)
T=$1; S=$2; LOG=${3:-/tmp/gpurun_last.log}
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then break; fi
  sleep 120
done
echo "gpurun rc=$rc" >> $LOG
