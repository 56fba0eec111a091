#!/bin/bash
# usage: gpurun_retry.sh <timeout> [--gpus N] -- <command>: retries while the pod answers "busy" (exit 3, nothing charged)
for i in $(seq 1 20); do
  gpurun --timeout "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
