#!/bin/bash
# 8 GPUs: TP parity at world 8 (step kernel + multi-launch, 7B per-rank shapes), bench at N=8 for both decode paths,
# and BASELINE's Llama-2-70B Q4 GQA configuration at tp8
mkdir -p gpurun_out
IFB_TP_7B=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29521 tests/tp_worker.py > gpurun_out/r2tp8_worker.log 2>&1
echo "tp_worker world 8 rc=$?"; grep "ok\|TP_WORKER\|rror" gpurun_out/r2tp8_worker.log | tail -12
show() {
python -c "
import json,sys
try:
    d=json.load(open('$1')); print(d['n_gpus'], d['config']['workload'][:24], d['decode_path'], round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['step']['frac'],3), 'verify err', d['verify']['max_abs_err'], 'tok', d['verify']['token'], d['verify']['reference_token'])
except Exception as e: print('no json', e)
"
}
for st in 1 0; do
  f=gpurun_out/r2tp8_bench_n8_step${st}
  IFB_STEP=$st timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 2954$st bench.py --gpus 8 --steps 100 --warmup 5 > $f.json 2> $f.err
  echo "bench 7B N=8 step=$st rc=$?"; show $f.json
  grep -v "Warn\|warn\|return func\|^$\|\*\*\*\|OMP_NUM\|NCCL version" $f.err | tail -3
done
f=gpurun_out/r2tp8_bench70b_n8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 60 --warmup 5 --workload llama2_70b_q4 > $f.json 2> $f.err
echo "bench 70B N=8 rc=$?"; show $f.json
grep -v "Warn\|warn\|return func\|^$\|\*\*\*\|OMP_NUM\|NCCL version" $f.err | tail -5
