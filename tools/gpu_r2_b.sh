#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 5 240 python tools/diag_fused.py > gpurun_out/diag_fused.log 2>&1; echo "diag_fused rc=$?"
grep -E "DIAG_FUSED|TIMING|stage out|Error|error" gpurun_out/diag_fused.log | head -40
if ! grep -q "DIAG_FUSED ALL OK" gpurun_out/diag_fused.log; then
  echo "!! fused kernel unhealthy: running the rest with PIFPAF_FUSE_DW=0"; export PIFPAF_FUSE_DW=0
fi
timeout -k 5 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --overlap 0 --quick --dump-ops gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.log
