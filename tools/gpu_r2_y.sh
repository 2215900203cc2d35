#!/bin/bash
# round-2 session y: ncu --set full of the one-CTA GEMM (stage 2) and the depthwise kernels on the final tree
set -u
mkdir -p gpurun_out
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k 'regex:^k_gemm_tc$|k_dwconv5_tma' -s 26 -c 9 \
   -o gpurun_out/prof_stage2 -f python tools/prof_target.py net 0 > gpurun_out/ncu_full_stage2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_stage2.log
ls -la gpurun_out/prof_stage2.ncu-rep
