#!/bin/bash
# round-2 session p: pair GEMM as the default (classes 1 + 2 + 8), heads class, whole GPU test suite, bench
set -u
mkdir -p gpurun_out
for s in 0 11 27 11 0; do
  PIFPAF_GEMM_PAIR=$s timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2
done
PIFPAF_GEMM_PAIR=27 timeout -k 5 300 python -m pytest tests/test_network_gpu.py -m gpu -q -x > gpurun_out/pytest_p1.log 2>&1; echo "pytest network (heads on pairs) rc=$?"; tail -3 gpurun_out/pytest_p1.log
timeout -k 5 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_p2.log 2>&1; echo "pytest all gpu rc=$?"; tail -5 gpurun_out/pytest_p2.log
DIAG_BATCH=8 PIFPAF_GEMM_PAIR=0 timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2 | head -1
DIAG_BATCH=8 PIFPAF_GEMM_PAIR=11 timeout -k 5 120 python tools/diag_perop.py 2>&1 | grep -v Warning | tail -2 | head -1
timeout -k 5 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['by_kind_ms'], d['roofline'].get('forward_ms'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_p.err').read()[-2000:])
PY
