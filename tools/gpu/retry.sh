#!/bin/bash
# tools/gpu/retry.sh [gpurun options ...] -- <command>: retry while the pod answers "busy" (exit code 3, nothing charged)
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
