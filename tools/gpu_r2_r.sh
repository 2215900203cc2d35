#!/bin/bash
# round-2 session r: 64-channel pitch of the depthwise INPUT tensors only; SMs reserved for the overlapped decode
set -u
mkdir -p gpurun_out
for v in 16 64 16 64; do
  PIFPAF_DWIN_PAD=$v timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
for r in 18 10 6 3 0; do
  echo "== reserve $r SMs"; PIFPAF_RESERVE_SMS=$r timeout -k 5 200 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['forward_ms'], d['decoder_only']['ms_per_batch'])"
done
PIFPAF_DWIN_PAD=64 timeout -k 5 400 python -m pytest tests/test_network_gpu.py -m gpu -q -x > gpurun_out/pytest_r.log 2>&1; echo "pytest network rc=$?"; tail -3 gpurun_out/pytest_r.log
