#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers busy / transient
t=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
