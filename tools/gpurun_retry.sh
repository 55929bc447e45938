#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged):
#   tools/gpurun_retry.sh <log> <timeout-seconds> '<command>'
log=$1; tmo=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $tmo -- "$@" > $log 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
