#!/bin/bash
# Usage: tools/gpurun_retry_n.sh <gpus> <timeout-seconds> '<command>'
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i: no GPU slot, sleeping 150 s"
  sleep 150
done
exit 3
