#!/bin/bash
# Local helper (CPU box): retry a gpurun call while the pod answers "busy" (exit 3); everything else is returned as-is.
#   tools/gpurun_retry.sh [--gpus N] --timeout S -- '<command>'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
