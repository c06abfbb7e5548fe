#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>  -- retries while the pod answers "transient" / busy (rc 3), up to 12 times
LOG=$1; shift
for i in $(seq 1 12); do
  gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|rc=3\|no box" "$LOG"; then sleep 120; continue; fi
  break
done
