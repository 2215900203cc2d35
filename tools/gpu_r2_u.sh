#!/bin/bash
# round-2 session u: what do the scattered pieces of the bins layout cost -- the same GEMMs writing one contiguous row
set -u
mkdir -p gpurun_out
for d in 0 8 0 8; do
  PIFPAF_GEMM_DEBUG=$d timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
