#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3): scripts/gpurun_retry.sh <timeout> <cmd...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
