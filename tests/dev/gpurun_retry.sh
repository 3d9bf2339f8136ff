#!/bin/bash
# usage: gpurun_retry.sh <timeout> '<command>'  — retries while the pod answers busy / another call is still registered;
# log in gpurun_out/retry.log
mkdir -p gpurun_out
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > gpurun_out/retry.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/retry.log || [ $rc -eq 3 ]; then sleep 8; continue; fi
  if [ $rc -eq 2 ] && grep -qi "another call\|in flight\|running" gpurun_out/retry.log; then sleep 20; continue; fi
  exit $rc
done
exit 3
