#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit 3)
log="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$log"; exit 3
