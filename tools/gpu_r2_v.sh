#!/bin/bash
# round-2 session v: bins pieces in multiples of 32 channels (64-byte bursts) instead of 16
set -u
mkdir -p gpurun_out
for g in 16 32 16 32; do
  PIFPAF_BIN_PAD=$g timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
