#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout_s> <command...>   retries while the pod answers busy (exit 3), nothing is charged then
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
