#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout-seconds> '<command>'   - retries while the pod answers "transient" (nothing charged)
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
