#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3 / "transient"); usage: tools/gpurun_retry.sh TIMEOUT 'command' [extra gpurun args]
T=$1; CMD=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
tail -40 /tmp/gpurun_last.log
exit $rc
