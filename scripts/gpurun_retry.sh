#!/bin/bash
# gpurun_retry.sh <timeout> [--gpus N] -- '<command>': retries while the pool answers "busy" (exit 3: nothing charged)
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
