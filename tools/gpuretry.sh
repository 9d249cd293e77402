#!/bin/bash
# usage: gpuretry.sh <timeout> <cmd...>  — retries while the pod reports "busy/draining" (rc 3)
T=$1; shift
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
