#!/bin/bash
# 4 GPUs: TP parity at world 4 (step kernel + multi-launch), bench at N=2 and N=4 for both decode paths
mkdir -p gpurun_out
IFB_TP_7B=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29521 tests/tp_worker.py > gpurun_out/r2tp4_worker.log 2>&1
echo "tp_worker world 4 rc=$?"; grep "ok\|TP_WORKER\|rror" gpurun_out/r2tp4_worker.log | tail -12
for n in 2 4; do
  for st in 1 0; do
    IFB_STEP=$st timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2tp4_bench_n${n}_step${st}.json 2> gpurun_out/r2tp4_bench_n${n}_step${st}.err
    echo "bench N=$n step=$st rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2tp4_bench_n${n}_step${st}.json')); print(d['n_gpus'], d['decode_path'], round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'verify err', d['verify']['max_abs_err'], 'tok', d['verify']['token'], d['verify']['reference_token'])
except Exception as e: print('no json', e)
"
    grep -v "Warn\|warn\|return func\|^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r2tp4_bench_n${n}_step${st}.err | tail -3
  done
done
