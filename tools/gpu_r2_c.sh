#!/bin/bash
# round-2 session c: all GPU tests, smoke, the full bench line (extra configs, library + CPU baselines), reference arm
set -u
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
timeout -k 5 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -c 1200 gpurun_out/bench_ref.log
