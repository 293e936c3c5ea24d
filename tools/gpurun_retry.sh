#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh <gpurun args...>
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
