#!/bin/bash
# round-2 validation session: GPU tests, smoke, bench (overlap on/off), reference arm
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --overlap 0 --quick > gpurun_out/bench_no_overlap.log 2>&1; echo "bench2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_ref.log
