#!/bin/bash
# round-2 last check after the decoder changes: whole GPU suite, smoke, quick bench
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final2.log
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final2.log
timeout -k 5 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2_final2_bench_quick.json 2>/dev/null; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_final2_bench_quick.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['by_kind_ms'], d['decoder_only']['ms_per_batch'], d['clocks'])
PY
