#!/bin/bash
# 2-GPU bench exactly as the driver launches it (torchrun, one rank per GPU), both arms
set +e
mkdir -p gpurun_out
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_n2.log | cut -c1-400
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
