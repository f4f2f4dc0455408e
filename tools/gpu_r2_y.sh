#!/bin/bash
# the persistent step kernel under tensor parallelism after the last edits of ifb_step.cu (default path from 4 GPUs on)
mkdir -p gpurun_out
IFB_STEP=1 IFB_TP_7B=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29571 tests/tp_worker.py > gpurun_out/r2y_worker_step.log 2>&1
echo "tp_worker world 2 IFB_STEP=1 rc=$?"; grep "ok\|TP_WORKER\|rror\|ssert" gpurun_out/r2y_worker_step.log | tail -10 | cut -c1-200
IFB_STEP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r2y_bench_n2_step1.json 2> gpurun_out/r2y_bench_n2_step1.err
echo "bench N=2 step kernel rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench_n2_step1.json')); print(d['value'], d['decode_path'], d['verify'])"
