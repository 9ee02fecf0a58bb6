#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <ngpus> <timeout-seconds> <script> [logfile]  -- multi-GPU box; retries while the pod answers "busy"
N=$1; T=$2; S=$3; LOG=${4:-/tmp/gpurun_last.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "bash $S" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then break; fi
  sleep 150
done
echo "gpurun rc=$rc" >> $LOG
