#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy" (rc 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
