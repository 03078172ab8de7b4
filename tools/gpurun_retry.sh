#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT LOGFILE COMMAND...  -- repeats a gpurun call while the pod answers "busy" (exit 3)
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
