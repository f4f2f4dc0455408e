#!/bin/bash
# 2 GPUs: tensor-parallel parity (toy + 7B per-rank shapes, both decode paths) and bench at N=2 with both paths
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2tp2_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/tp_worker.py > gpurun_out/r2tp2_worker.log 2>&1
echo "tp_worker rc=$?"; tail -25 gpurun_out/r2tp2_worker.log
IFB_STEP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 tests/tp_worker.py > gpurun_out/r2tp2_worker_multi.log 2>&1
echo "tp_worker (multi-launch) rc=$?"; tail -8 gpurun_out/r2tp2_worker_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2tp2_bench_step.json 2> gpurun_out/r2tp2_bench_step.err
echo "bench N=2 step rc=$?"; cat gpurun_out/r2tp2_bench_step.json; tail -3 gpurun_out/r2tp2_bench_step.err
IFB_STEP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2tp2_bench_multi.json 2> gpurun_out/r2tp2_bench_multi.err
echo "bench N=2 multi rc=$?"; cat gpurun_out/r2tp2_bench_multi.json; tail -3 gpurun_out/r2tp2_bench_multi.err
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2tp2_bench_n1.json 2> gpurun_out/r2tp2_bench_n1.err
echo "bench N=1 rc=$?"; cat gpurun_out/r2tp2_bench_n1.json; tail -3 gpurun_out/r2tp2_bench_n1.err
