#!/bin/bash
# round-2 session i: k_grow phase clocks (rounds, seeds grown, clocks per phase) on the decoder workloads
set -u
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_i.log
for k in 6 0; do
  echo "== PIFPAF_GROW_DEFER=$k"; PIFPAF_GROW_DEFER=$k timeout -k 5 200 python tools/diag_decoder_perf.py 2>&1 | tail -8
done
