#!/bin/bash
# round-2 session l: weights-resident GEMMs with more A stages (narrower n blocks)
set -u
mkdir -p gpurun_out
for s in 0 6 7 8 0 8; do
  PIFPAF_GEMM_RES_STAGES=$s timeout -k 5 200 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
PIFPAF_GEMM_RES_STAGES=8 timeout -k 5 600 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "every_op or full_size" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_l.log
