#!/bin/bash
# round-2 session e: PDL and channel-block-fastest stride-2 depthwise: correctness (all GPU tests) and A/B timing
set -u
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_e.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  for b in 64 8; do
    echo "== PDL=$1 DW_CBF=$2 batch=$b"
    PIFPAF_PDL=$1 PIFPAF_DW_CBF=$2 timeout -k 5 300 python bench.py --steps 20 --warmup 3 --quick --batch $b > gpurun_out/bench_ab.log 2>&1
    python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_ab.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_ab.log').read()[-800:])
else:
    d=json.loads(l[-1]); r=d['roofline']
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'fwd_ms(timed pass)', r['forward_ms'], r['by_kind_ms'])
PY
  done
done
