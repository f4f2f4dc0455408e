#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2tp2_bench_step.json 2> gpurun_out/r2tp2_bench_step.err
echo "bench N=2 step rc=$?"; cat gpurun_out/r2tp2_bench_step.json; grep -v "Warn\|warn\|return func\|^$" gpurun_out/r2tp2_bench_step.err | tail -5
