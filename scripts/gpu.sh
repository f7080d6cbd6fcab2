#!/bin/bash
# scripts/gpu.sh <log> <timeout-seconds> <command...>: gpurun with retries while the pod answers "busy" (exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
exit 3
