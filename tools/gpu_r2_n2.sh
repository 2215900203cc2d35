#!/bin/bash
# 2 GPUs, launched like the driver does: strong-scaled C0 (32 images per GPU) + the weak point; reference arm under torchrun
set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/bench_n2.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','weak_scaling','annotations_last_step','gpu_launches')}, d['e2e'], d['config']['workload'])"
tail -3 gpurun_out/bench_n2.log | cut -c1-300
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"
grep '^{' gpurun_out/bench_ref_n2.log | tail -1 | cut -c1-400
