#!/bin/bash
# usage: tools/gpu_retry.sh <gpus> <timeout> <script>   -- retries while the pod answers "transient" (nothing charged)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun ${1:+--gpus $1} --timeout $2 -- "bash $3" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up"
