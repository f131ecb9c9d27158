#!/bin/sh
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy/transient (exit code 3 / status=transient)
LOG=$1; shift; TMO=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
