#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <log> <timeout_s> <command...>
N=$1; shift; LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
