#!/bin/bash
# round-2 session m: two-CTA clusters sharing A by TMA multicast in the weights-resident GEMMs with two n blocks
set -u
mkdir -p gpurun_out
echo "== small correctness first (hang guard)"
PIFPAF_GEMM_MC=1 timeout -k 5 150 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "every_op" > gpurun_out/pytest_m1.log 2>&1; echo "pytest every_op rc=$?"; tail -3 gpurun_out/pytest_m1.log
for s in 0 1 0 1; do
  PIFPAF_GEMM_MC=$s timeout -k 5 200 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
PIFPAF_GEMM_MC=1 timeout -k 5 400 python -m pytest tests/test_network_gpu.py -m gpu -q -x > gpurun_out/pytest_m2.log 2>&1; echo "pytest network rc=$?"; tail -3 gpurun_out/pytest_m2.log
