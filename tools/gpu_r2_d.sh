#!/bin/bash
# round-2 session d: decoder with deferred seed selection (tests + timing, radius sweep), fused kernel with staged
# depthwise weights, CifDet / plugin tests, quick bench
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x -k "decoder or cifdet or plugin or predictor or fused" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_d.log
for k in 12 0 6 20; do
  echo "== PIFPAF_GROW_DEFER=$k"; PIFPAF_GROW_DEFER=$k timeout -k 5 200 python tools/diag_decoder_perf.py 2>&1 | tail -4
done
timeout -k 5 300 python tools/diag_fused.py > gpurun_out/diag_fused.log 2>&1; echo "diag_fused rc=$?"
grep -E "DIAG_FUSED|TIMING" gpurun_out/diag_fused.log
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_quick.log
