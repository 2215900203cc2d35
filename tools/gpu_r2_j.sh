#!/bin/bash
# round-2 session j: depthwise 5x5 on the tensor cores -- descriptor variants, then per-op times
set -u
mkdir -p gpurun_out
echo "== stride 1 only, small"; DIAG_DW_TC=1 timeout -k 5 240 python tools/diag_dwtc.py small 2>&1 | grep -v Warning | tail -12
echo "== stride 1 + 2, small"; DIAG_DW_TC=3 timeout -k 5 240 python tools/diag_dwtc.py small 2>&1 | grep -v Warning | tail -12
for v in "12 0" "16 0" "16 1"; do set -- $v
  echo "== perf pwid=$1 bo=$2 (stride 1)"; DIAG_DW_TC=1 DIAG_PWID=$1 DIAG_BO=$2 timeout -k 5 200 python tools/diag_dwtc.py perf 2>&1 | grep -v Warning | tail -5
done
echo "== perf stride 1 + 2"; DIAG_DW_TC=3 timeout -k 5 200 python tools/diag_dwtc.py perf 2>&1 | grep -v Warning | tail -5
