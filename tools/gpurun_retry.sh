#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] '<command>'  -- retries while the pod answers "no slot" (exit 3)
t=$1; shift
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t "${extra[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
