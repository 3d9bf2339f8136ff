#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <timeout> '<command>'
mkdir -p gpurun_out
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3" > gpurun_out/retry_n.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/retry_n.log || [ $rc -eq 3 ]; then sleep 8; continue; fi
  if [ $rc -eq 2 ] && grep -qi "another call\|in flight\|running" gpurun_out/retry_n.log; then sleep 20; continue; fi
  exit $rc
done
exit 3
