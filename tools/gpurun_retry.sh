#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged): tools/gpurun_retry.sh LOGFILE gpurun-args...
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
exit 3
