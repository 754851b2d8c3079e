#!/bin/bash
# usage: scripts/gpurun_retry.sh TIMEOUT_S [--gpus N] -- 'command'   (retries while the pod answers "busy / transient")
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 40 tries"; exit 3
