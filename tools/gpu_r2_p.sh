#!/bin/bash
# warp-private pipelines of the step kernel: parity first, then A/B timing on the same box
mkdir -p gpurun_out
IFB_STEP_PRIV=1 timeout 900 python -m pytest tests/test_gpu_step.py -x -q > gpurun_out/r2p_step.log 2>&1
echo "step tests rc=$?"; tail -4 gpurun_out/r2p_step.log
: > gpurun_out/r2p_sweep.txt
IFB_STEP=1 IFB_STEP_PRIV=1 timeout 200 python tools/step_sweep.py "step_private" 60 >> gpurun_out/r2p_sweep.txt 2>&1
IFB_STEP=1 IFB_STEP_PRIV=0 timeout 200 python tools/step_sweep.py "step_shared_ring" 60 >> gpurun_out/r2p_sweep.txt 2>&1
IFB_STEP=0 timeout 200 python tools/step_sweep.py "multi_launch" 60 >> gpurun_out/r2p_sweep.txt 2>&1
grep -v "^$" gpurun_out/r2p_sweep.txt | tail -12
IFB_STEP=1 IFB_STEP_PRIV=1 timeout 200 python tools/trace_step2.py > gpurun_out/r2p_trace.txt 2>&1
tail -30 gpurun_out/r2p_trace.txt
