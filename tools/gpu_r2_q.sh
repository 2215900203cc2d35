#!/bin/bash
# round-2 session q: pair GEMM default -- bitwise test, launch list, DRAM traffic of one step, ncu --set full of the
# pair kernel, full bench line
set -u
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "pair_gemm" > gpurun_out/pytest_q.log 2>&1; echo "pytest pair bitwise rc=$?"; tail -12 gpurun_out/pytest_q.log
echo "== ncu launch list of the bench command"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_q_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu dram traffic of one bench step (57 network launches)"
timeout -k 5 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:'k_gemm_tc|k_dwconv5|k_input_conv' -s 171 -c 57 --csv --log-file gpurun_out/r2_q_dram_traffic_step.csv \
   python bench.py --steps 1 --warmup 3 --quick > gpurun_out/bench_under_ncu_traffic.log 2>&1; echo "rc=$?"
python tools/summarize_traffic.py gpurun_out/r2_q_dram_traffic_step.csv gpurun_out/r2_dram_traffic_bench_step.json
echo "== ncu full: pair kernel (stage 3 resident, stage 4 streaming, conv5)"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc2' -s 40 -c 14 \
   -o gpurun_out/prof_pair -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_pair.log 2>&1; echo "rc=$?"
echo "== bench (full line)"
timeout -k 5 900 python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r2_q_per_op_bs64.json > gpurun_out/r2_q_bench_n1_full.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_q_bench_n1_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['by_kind_ms'], d['roofline']['traffic'])
print({k: (v.get('images_per_s'), v.get('forward_ms')) for k, v in d['extra_configs'].items()})
print(d['library_baseline'])
PY
ls -la gpurun_out/
