#!/bin/bash
# gpurun with retries while the pod answers "transient"/busy (nothing is charged for those):  tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
