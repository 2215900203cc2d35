#!/bin/bash
# round-2 session x: scatter pieces written as their own [pixels][width] tensors (debug 16) against one contiguous row (8)
set -u
mkdir -p gpurun_out
for d in 0 16 8 16; do
  PIFPAF_GEMM_DEBUG=$d timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
