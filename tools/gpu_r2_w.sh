#!/bin/bash
# round-2 session w: k_grow with the destination components and the list lengths staged in shared memory
set -u
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_decoder_gpu.py tests/test_cifdet_gpu.py -m gpu -q -x > gpurun_out/pytest_w.log 2>&1; echo "pytest decoder rc=$?"; tail -3 gpurun_out/pytest_w.log
timeout -k 5 200 python tools/diag_decoder_perf.py 2>&1 | tail -12
