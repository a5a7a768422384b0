#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh <out-file> <gpurun args...>
out=$1; shift
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy" >> "$out.retries"
  sleep 60
done
exit 3
