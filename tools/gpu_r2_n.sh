#!/bin/bash
# round-2 session n: where does a GEMM launch spend its time -- the same forward without epilogue stores (1),
# without MMAs (2), without TMEM reads (4) and combinations (results are wrong, timing only)
set -u
mkdir -p gpurun_out
for d in 0 1 2 4 3 7 0; do
  PIFPAF_GEMM_DEBUG=$d timeout -k 5 200 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
