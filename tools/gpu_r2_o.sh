#!/bin/bash
# round-2 session o: CTA-pair GEMM (k_gemm_tc2, cta_group::2) -- correctness under a hang guard, then per-op times
set -u
mkdir -p gpurun_out
echo "== every_op with all GEMM classes on the pair kernel"
PIFPAF_GEMM_PAIR=7 timeout -k 5 150 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "every_op" > gpurun_out/pytest_o1.log 2>&1; rc=$?; echo "pytest every_op rc=$rc"; tail -15 gpurun_out/pytest_o1.log
if [ $rc -ne 0 ]; then
  for m in 1 2 4; do
    PIFPAF_GEMM_PAIR=$m timeout -k 5 100 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "every_op and bins-False" > gpurun_out/pytest_o1_$m.log 2>&1; echo "mask $m rc=$?"; tail -4 gpurun_out/pytest_o1_$m.log
  done
fi
for s in 0 7 0 1 2 3; do
  PIFPAF_GEMM_PAIR=$s timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
PIFPAF_GEMM_PAIR=7 timeout -k 5 400 python -m pytest tests/test_network_gpu.py -m gpu -q -x > gpurun_out/pytest_o2.log 2>&1; echo "pytest network rc=$?"; tail -3 gpurun_out/pytest_o2.log
