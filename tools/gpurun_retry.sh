#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command string>   -- retries while the pod is busy (rc 3)
LOG=$1; TO=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc attempt=$attempt" >> $LOG; break; fi
  sleep 45
done
echo finished >> $LOG
