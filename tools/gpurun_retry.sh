#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- <command>   (retries while the pod answers busy / transient)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$OUT" | tail -6
  if echo "$OUT" | grep -q "status=transient\|answers busy\|rc=3\|no box"; then sleep 90; continue; fi
  break
done
