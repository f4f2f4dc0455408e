#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 0"; do
  set -- $cfg
  IFB_STEP_BATCH=$1 IFB_STEP_INFLIGHT=$2 timeout 200 python tools/trace_step2.py > gpurun_out/r2c_trace_b$1_i$2.txt 2>&1
  cat gpurun_out/r2c_trace_b$1_i$2.txt
done
