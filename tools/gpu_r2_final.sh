#!/bin/bash
# round-2 final check of the tree as committed: whole GPU suite, smoke, launch list + DRAM traffic of the bench command,
# full bench line, reference arm
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/pytest_gpu_final.log
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout -k 5 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:'k_gemm_tc|k_dwconv5|k_input_conv' -s 171 -c 57 --csv --log-file gpurun_out/r2_final_dram_traffic_step.csv \
   python bench.py --steps 1 --warmup 3 --quick > gpurun_out/bench_under_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python tools/summarize_traffic.py gpurun_out/r2_final_dram_traffic_step.csv gpurun_out/r2_dram_traffic_bench_step.json
cp gpurun_out/r2_dram_traffic_bench_step.json profiles/r2_dram_traffic_bench_step.json
timeout -k 5 900 python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r2_final_per_op_bs64.json > gpurun_out/r2_final_bench_n1_full.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_final_bench_n1_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['by_kind_ms'], d['roofline']['traffic'], d['clocks'])
print({k: (v.get('images_per_s'), v.get('forward_ms')) for k, v in d['extra_configs'].items()}, d['cpu_baseline']['value'])
PY
timeout -k 5 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"; tail -c 300 gpurun_out/r2_final_bench_reference_arm.json
