#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py -x -q --timeout 600 > gpurun_out/r2j_step.log 2>&1
echo "step tests rc=$?" >> gpurun_out/r2j_step.log; tail -15 gpurun_out/r2j_step.log
: > gpurun_out/r2j_sweep.txt
for cfg in "1 0" "2 0"; do
  set -- $cfg
  IFB_STEP=1 IFB_STEP_BATCH=$1 IFB_STEP_INFLIGHT=$2 timeout 120 python tools/step_sweep.py "batch$1_inflight$2" >> gpurun_out/r2j_sweep.txt 2>&1
done
IFB_STEP=0 timeout 120 python tools/step_sweep.py multi_launch >> gpurun_out/r2j_sweep.txt 2>&1
cat gpurun_out/r2j_sweep.txt
IFB_STEP=1 IFB_STEP_BATCH=1 timeout 200 python tools/trace_step2.py > gpurun_out/r2j_trace.txt 2>&1
cat gpurun_out/r2j_trace.txt
