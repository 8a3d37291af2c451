#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: gpurun_retry.sh <timeout> <log> [--gpus N] -- <command>
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
