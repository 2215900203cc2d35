#!/bin/bash
# Short gpurun call for kernel iterations: GEMM self-check -> GPU tests -> per-kind forward timing -> one bench line.  Logs land in gpurun_out/.
set +e
mkdir -p gpurun_out
echo "== diag_gemm"; timeout -k 5 150 python tools/diag_gemm.py > gpurun_out/diag_gemm.log 2>&1; echo "rc=$?"
if ! grep -q "DIAG_GEMM ALL OK" gpurun_out/diag_gemm.log; then
  echo "!! GEMM self-check failed or hung: aborting early"; tail -30 gpurun_out/diag_gemm.log; exit 0
fi
echo "== pytest -m gpu (network + a decoder subset)"
timeout -k 5 300 python -m pytest tests/test_network_gpu.py -m gpu -q -x --durations=4 -k "not resnet" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/pytest_gpu.log
echo "== diag_net"
DIAG_NET_FAST=1 timeout -k 5 300 python tests/diag/diag_net.py > gpurun_out/diag_net.log 2>&1; echo "rc=$?"
grep -E "forward bs64|input_conv:|gemm_tc:|dwconv:|DIAG_NET|BAD|rror" gpurun_out/diag_net.log | head -14
echo "== bench"; timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/bench.log | cut -c1-600
