#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3, nothing charged):
#   [GPUS=N] tools/gpurun_retry.sh <timeout> <logfile> <command...>
t=$1; log=$2; shift 2
extra=""
if [ -n "${GPUS:-}" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun $extra --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 8
done
exit 3
