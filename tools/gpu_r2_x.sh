#!/bin/bash
mkdir -p gpurun_out
IFB_TP_7B=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 tests/tp_worker.py > gpurun_out/r2x_worker.log 2>&1
echo "tp_worker world 2 rc=$?"; grep "moe\|ok\|TP_WORKER\|rror\|ssert" gpurun_out/r2x_worker.log | tail -16 | cut -c1-260
