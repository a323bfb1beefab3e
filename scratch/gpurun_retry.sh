#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / transient); usage: gpurun_retry.sh <timeout> '<command>'
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then sleep 100; continue; fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
