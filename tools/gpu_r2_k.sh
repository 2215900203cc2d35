#!/bin/bash
# round-2 session k: FFMA2 depthwise kernel -- bitwise tests, per-op times
set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "every_op or bitwise or full_size or linearity" > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_k.log
DIAG_DW_TC=0 timeout -k 5 200 python tools/diag_dwtc.py perf 2>&1 | grep -v Warning | tail -5
