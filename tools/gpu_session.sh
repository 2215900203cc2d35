#!/bin/bash
# One gpurun call: diagnostics -> GPU tests -> bench -> ncu.  Logs land in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag_gemm"; timeout 420 python tools/diag_gemm.py > gpurun_out/diag_gemm.log 2>&1; echo "rc=$?"
grep -E "DIAG_GEMM|BAD|EXCEPTION" gpurun_out/diag_gemm.log | head -20
if grep -q "\[tcgen05\].*\(BAD\|EXCEPTION\)" gpurun_out/diag_gemm.log || ! grep -q "DIAG_GEMM" gpurun_out/diag_gemm.log; then
  echo "!! tcgen05 GEMM not healthy: running decoder-only tests"
  tail -40 gpurun_out/diag_gemm.log
  timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -15 gpurun_out/pytest_gpu.log
  exit 0
fi
echo "== diag_net"; timeout 600 python tools/diag_net.py > gpurun_out/diag_net.log 2>&1; echo "rc=$?"
tail -45 gpurun_out/diag_net.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/bench.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/bench_ref.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full: gemm, cifhr, dwconv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_cifhr_tiles|k_dwconv|k_grow' -s 80 -c 12 \
   -o gpurun_out/prof_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
