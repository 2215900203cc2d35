#!/bin/bash
# round-2 session g: fused dw->GEMM kernel with 16 depthwise warps (setmaxnreg): bit-exactness, timing, ncu
set -u
mkdir -p gpurun_out
timeout -k 5 90 python tools/diag_fused_small.py > gpurun_out/diag_fused_small.log 2>&1; rc=$?; echo "diag_fused_small rc=$rc"
tail -3 gpurun_out/diag_fused_small.log
if [ $rc -ne 0 ]; then echo "!! fused kernel hangs or fails: stop"; exit 0; fi
timeout -k 5 240 python tools/diag_fused.py > gpurun_out/diag_fused.log 2>&1; echo "diag_fused rc=$?"
grep -E "DIAG_FUSED|TIMING|Error|error" gpurun_out/diag_fused.log | head -20
grep -E "op[0-9]+ (dw_conv1x1|dwconv)" gpurun_out/diag_fused.log | head -40
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:'k_dw_gemm' -s 12 -c 2 \
   -o gpurun_out/prof_fused16 -f python tools/prof_target.py net 1 > gpurun_out/ncu_full_fused16.log 2>&1; echo "ncu fused rc=$?"
