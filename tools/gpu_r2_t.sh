#!/bin/bash
# round-2 session t: tracking plugin route on the GPU, resident weight loads ahead of the PDL wait (8 / 64 images), all tests
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_t.log 2>&1; echo "pytest all gpu rc=$?"; tail -6 gpurun_out/pytest_t.log
cat gpurun_out/plugin_gpu_report.txt 2>/dev/null | tail -2
for b in 8 64; do
  echo "== batch $b"; timeout -k 5 200 python bench.py --batch $b --steps 30 --warmup 5 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['forward_ms'], d['roofline']['by_kind_ms'], d['decoder_only']['ms_per_batch'])"
done
