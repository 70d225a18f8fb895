#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3 / "transient"):  tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|no box or slot" "$log"; then sleep 120; continue; fi
  exit $rc
done
exit 3
