#!/bin/bash
# tools/gpurun_retry.sh <timeout> <outfile> <command...>: retry while the pod answers busy (exit 3)
to=$1; out=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
