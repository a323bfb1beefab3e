#!/bin/bash
# gpurun --gpus N with retries; usage: gpurun_retry_n.sh <gpus> <timeout> '<command>'
for i in $(seq 1 15); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|rc=3"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
