#!/bin/bash
# last check of the tree as committed: whole GPU suite + smoke (+ the raw-image bench variant)
set +e
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/pytest_gpu_final.log
timeout -k 5 150 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout -k 5 200 python bench.py --raw-input --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_raw_input.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_raw_input.log | cut -c1-300
