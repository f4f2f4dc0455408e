#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py -x -q --timeout 600 > gpurun_out/r2h_step.log 2>&1
echo "step tests rc=$?" >> gpurun_out/r2h_step.log; tail -15 gpurun_out/r2h_step.log
: > gpurun_out/r2h_sweep.txt
IFB_STEP_LOCAL_LL=0 IFB_STEP_BATCH=1 timeout 120 python tools/step_sweep.py "ll0_batch1" >> gpurun_out/r2h_sweep.txt 2>&1
IFB_STEP_LOCAL_LL=1 IFB_STEP_BATCH=1 timeout 120 python tools/step_sweep.py "ll1_batch1" >> gpurun_out/r2h_sweep.txt 2>&1
IFB_STEP_LOCAL_LL=1 IFB_STEP_BATCH=2 timeout 120 python tools/step_sweep.py "ll1_batch2" >> gpurun_out/r2h_sweep.txt 2>&1
cat gpurun_out/r2h_sweep.txt
IFB_STEP_BATCH=1 timeout 200 python tools/trace_step2.py > gpurun_out/r2h_trace.txt 2>&1
cat gpurun_out/r2h_trace.txt
