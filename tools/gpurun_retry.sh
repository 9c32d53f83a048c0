#!/bin/bash
# gpurun with retries while the pod answers "busy" (status transient / rc 3); usage: gpurun_retry.sh <timeout> <command...>
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "gpurun: still busy after 30 attempts"; exit 3
