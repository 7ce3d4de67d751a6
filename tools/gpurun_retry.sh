#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged): tools/gpurun_retry.sh <timeout_s> '<command>' <logfile>
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2" > "$3" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
