#!/bin/bash
# round-2: last check of the committed tree (after the debug-only change to the scatter table): all GPU tests, smoke
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final3.log
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke_final3.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final3.log
