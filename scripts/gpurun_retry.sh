#!/bin/bash
# development tool: gpurun with retries while the pod answers "busy" (exit 3: nothing charged)
#   scripts/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
