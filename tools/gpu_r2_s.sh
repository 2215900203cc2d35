#!/bin/bash
# round-2 session s: stride-2 depthwise input pitch (producer writes whole rows); SM reservation at 8 / 16 / 32 images
set -u
mkdir -p gpurun_out
for v in 16 64 16 64; do
  PIFPAF_DWIN_PAD=$v timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2 | cut -c1-330
done
for b in 8 16 32; do for r in 0 4 8 18; do
  echo "== batch $b reserve $r"; PIFPAF_RESERVE_SMS=$r timeout -k 5 200 python bench.py --batch $b --steps 30 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['forward_ms'], d['decoder_only']['ms_per_batch'])"
done; done
timeout -k 5 400 python -m pytest tests/test_network_gpu.py -m gpu -q -x > gpurun_out/pytest_s.log 2>&1; echo "pytest network rc=$?"; tail -3 gpurun_out/pytest_s.log
