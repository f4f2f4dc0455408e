#!/bin/bash
# round 2, call B: step kernel parity + knob sweep + trace
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py -x -q --timeout 600 > gpurun_out/r2b_step.log 2>&1
echo "step tests rc=$?" >> gpurun_out/r2b_step.log
tail -40 gpurun_out/r2b_step.log
: > gpurun_out/r2b_sweep.txt
for cfg in "1 0" "2 0" "4 0" "1 2" "1 3" "1 4" "1 6" "2 3" "4 3"; do
  set -- $cfg
  IFB_STEP_BATCH=$1 IFB_STEP_INFLIGHT=$2 timeout 120 python tools/step_sweep.py "batch$1_inflight$2" >> gpurun_out/r2b_sweep.txt 2>&1
done
IFB_STEP=0 timeout 120 python tools/step_sweep.py multi_launch >> gpurun_out/r2b_sweep.txt 2>&1
cat gpurun_out/r2b_sweep.txt
IFB_STEP_BATCH=1 IFB_STEP_INFLIGHT=0 timeout 200 python tools/trace_step2.py > gpurun_out/r2b_trace_b1_i0.txt 2>&1
IFB_STEP_BATCH=1 IFB_STEP_INFLIGHT=3 timeout 200 python tools/trace_step2.py > gpurun_out/r2b_trace_b1_i3.txt 2>&1
cat gpurun_out/r2b_trace_b1_i0.txt; cat gpurun_out/r2b_trace_b1_i3.txt
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 --ignore tests/test_gpu_step.py > gpurun_out/r2b_all.log 2>&1
echo "all gpu tests rc=$?"; tail -15 gpurun_out/r2b_all.log
