#!/bin/bash
# One gpurun call: GEMM self-check (abort early if unhealthy) -> GPU tests -> diagnostics -> bench -> ncu.
# Every step has a tight timeout; logs land in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag_gemm"; timeout -k 5 150 python tools/diag_gemm.py > gpurun_out/diag_gemm.log 2>&1; echo "rc=$?"
grep -E "DIAG_GEMM|BAD|EXCEPTION" gpurun_out/diag_gemm.log | head -20
if ! grep -q "DIAG_GEMM ALL OK" gpurun_out/diag_gemm.log; then
  echo "!! GEMM self-check failed or hung: aborting the session early"
  tail -30 gpurun_out/diag_gemm.log
  exit 0
fi
echo "== pytest -m gpu (all but resnet)"; timeout -k 5 400 python -m pytest tests -m gpu -q -x --durations=6 -k "not resnet" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_gpu.log
echo "== diag_net"; timeout -k 5 300 python tests/diag/diag_net.py > gpurun_out/diag_net.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/diag_net.log
echo "== decoder perf"; timeout -k 5 200 python tools/diag_decoder_perf.py > gpurun_out/decoder_perf.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/decoder_perf.log
echo "== smoke"; timeout -k 5 150 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout -k 5 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/bench.log
echo "== bench reference arm"; timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/bench_ref.log
echo "== pytest resnet"; timeout -k 5 300 python -m pytest tests -m gpu -q -k "resnet" > gpurun_out/pytest_resnet.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/pytest_resnet.log
echo "== ncu launch lists"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_decoder.csv \
   python tools/diag_decoder_perf.py 2 > gpurun_out/decoder_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu dram traffic of one bench step (57 network launches after 3 warm-up steps)"
timeout -k 5 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:'k_gemm_tc|k_dwconv5|k_input_conv' -s 171 -c 57 --csv --log-file gpurun_out/dram_traffic_step.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_traffic.log 2>&1; echo "rc=$?"
python tools/summarize_traffic.py gpurun_out/dram_traffic_step.csv gpurun_out/dram_traffic_step.json
echo "== ncu full"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_dwconv5' -s 70 -c 12 \
   -o gpurun_out/prof_net -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_net.log 2>&1; echo "rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:'k_grow|k_cifhr_tiles|k_nms|k_seed_sort|k_caf_scored' -s 40 -c 10 \
   -o gpurun_out/prof_dec -f python tools/diag_decoder_perf.py 2 > gpurun_out/ncu_full_dec.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
