#!/bin/bash
# round-2 profiling session: the two tests that failed, decoder timing, ncu of the fused dw->GEMM kernel and of k_grow
set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -q -k "three_hundred or reference_predictor" > gpurun_out/pytest_two.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_two.log
timeout -k 5 200 python tools/diag_decoder_perf.py > gpurun_out/decoder_perf.log 2>&1; echo "dec perf rc=$?"
cat gpurun_out/decoder_perf.log
# launch list of the decoder (per-kernel device time)
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_decoder.csv \
   python tools/prof_target.py dec > gpurun_out/dec_under_ncu.log 2>&1; echo "ncu dec list rc=$?"
# launch list of both forward schedules
timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_net_fused.csv \
   python tools/prof_target.py net 1 > gpurun_out/net1_under_ncu.log 2>&1; echo "ncu net fused list rc=$?"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_net_plain.csv \
   python tools/prof_target.py net 0 > gpurun_out/net0_under_ncu.log 2>&1; echo "ncu net plain list rc=$?"
# full captures: fused kernel (stage-2 and stage-3 launches of the second forward), k_grow
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:'k_dw_gemm' -s 12 -c 2 \
   -o gpurun_out/prof_fused -f python tools/prof_target.py net 1 > gpurun_out/ncu_full_fused.log 2>&1; echo "ncu fused rc=$?"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:'k_grow' -s 1 -c 3 \
   -o gpurun_out/prof_grow -f python tools/prof_target.py dec > gpurun_out/ncu_full_grow.log 2>&1; echo "ncu grow rc=$?"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:'k_dwconv5_tma' -s 19 -c 6 \
   -o gpurun_out/prof_dw -f python tools/prof_target.py net 0 > gpurun_out/ncu_full_dw.log 2>&1; echo "ncu dw rc=$?"
ls -la gpurun_out/
