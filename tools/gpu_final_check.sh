#!/bin/bash
# last check of the tree as committed: network GPU tests + the raw-image bench variant
set +e
mkdir -p gpurun_out
timeout -k 5 100 python -m pytest tests/test_network_gpu.py -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/pytest_gpu_final.log
timeout -k 5 100 python bench.py --raw-input --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_raw_input.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_raw_input.log | cut -c1-300
