#!/bin/bash
# gpurun_retry.sh <gpurun args...>: retry while the pod answers "busy" (exit code 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
