#!/bin/bash
# round-2 session h: pitch of the branch-internal tensors (multiple of 16 vs 64 channels): correctness + A/B timing
set -u
mkdir -p gpurun_out
PIFPAF_BRANCH_PAD=64 timeout -k 5 600 python -m pytest tests/test_network_gpu.py tests/test_accuracy_gpu.py -m gpu -q -x > gpurun_out/pytest_h.log 2>&1; echo "pytest(pad64) rc=$?"
tail -3 gpurun_out/pytest_h.log
for pad in 16 64 16 64; do
  echo "== PIFPAF_BRANCH_PAD=$pad"
  PIFPAF_BRANCH_PAD=$pad timeout -k 5 300 python bench.py --steps 20 --warmup 3 --quick --dump-ops gpurun_out/per_op_pad$pad.json > gpurun_out/bench_ab.log 2>&1
  python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_ab.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_ab.log').read()[-800:])
else:
    d=json.loads(l[-1]); r=d['roofline']
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'fwd_ms(timed pass)', r['forward_ms'], r['by_kind_ms'], 'gemm frac', r['frac'], r['per_bound']['frac_of_own_bound'])
PY
done
