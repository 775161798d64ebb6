#!/bin/bash
# gpurun_retry.sh <timeout> <logfile> <command...> — retries while the pod answers "busy" (exit 3, nothing charged).
t=$1; log=$2; shift 2
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
