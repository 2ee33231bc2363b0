#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3): tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
