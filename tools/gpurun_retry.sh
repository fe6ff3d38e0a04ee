#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh [gpurun flags ...] -- '<command>'   (retries while the pod answers "busy", exit code 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
