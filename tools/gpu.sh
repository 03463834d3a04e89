#!/bin/bash
# gpurun with retries while the pod has no free slot (gpurun exit code 3): tools/gpu.sh <timeout_s> '<command>'
# Run it in the FOREGROUND only: a detached retry loop re-submits the command and burns the GPU budget.
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  RC=$?
  if [ $RC -ne 3 ]; then exit $RC; fi
  sleep 60
done
exit 3
