#!/bin/bash
# usage: gpurun_retry.sh <timeout> '<command>'  — retries while the pod answers busy (exit 3); log in gpurun_out/retry.log
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > gpurun_out/retry.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/retry.log; then exit $rc; fi
  sleep 90
done
exit 3
