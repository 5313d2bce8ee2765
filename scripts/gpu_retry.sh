#!/bin/bash
# usage: gpu_retry.sh <outfile> <timeout> <command...>   — retries while the pod answers "busy" (exit 3)
out=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
