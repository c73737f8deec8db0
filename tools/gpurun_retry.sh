#!/bin/bash
# usage: tools/gpurun_retry.sh "<gpurun args>" '<command>'   -- retries while the pod answers busy / transient
ARGS="$1"; CMD="$2"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $ARGS -- "$CMD" > /tmp/gpurun_retry_$$.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_retry_$$.log || [ $rc -eq 3 ]; then
    echo "attempt $i: pod busy, retrying in 150 s"; sleep 150; continue
  fi
  break
done
tail -80 /tmp/gpurun_retry_$$.log
