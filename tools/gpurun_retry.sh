#!/bin/bash
# usage: [GPURUN_OPTS="--gpus 2"] tools/gpurun_retry.sh <tag> <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (nothing charged)
tag=$1; tmo=$2; shift 2
out=gpurun_out/gpurun_$tag.out
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_OPTS --timeout "$tmo" -- "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient" "$out" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  echo "gpurun rc=$rc attempt=$attempt" >> "$out"
  exit $rc
done
echo "gave up" >> "$out"
