#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> <timeout-seconds> '<command>'   -- retries while the pod answers "busy"
log="$1"; tmo="$2"; cmd="$3"
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun ${GPURUN_GPUS:+--gpus $GPURUN_GPUS} --timeout "$tmo" -- "$cmd" > "$log.tmp" 2>&1
  if grep -q "status=transient" "$log.tmp" || grep -q "rc=3" "$log.tmp"; then sleep 45; continue; fi
  break
done
mv "$log.tmp" "$log"
