#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2k_sweep.txt
# streaming + boundaries only (no task arithmetic; flag-in-data waits would never be satisfied -> counters mode)
IFB_STEP=1 IFB_STEP_LOCAL_LL=0 IFB_STEP_DEBUG=1 timeout 120 python tools/step_sweep.py "nocompute_counters" >> gpurun_out/r2k_sweep.txt 2>&1
IFB_STEP=1 IFB_STEP_LOCAL_LL=0 timeout 120 python tools/step_sweep.py "compute_counters" >> gpurun_out/r2k_sweep.txt 2>&1
cat gpurun_out/r2k_sweep.txt
sed -i 's/r2g_/r2k_/g' tools/gpu_r2_f.sh
IFB_STEP=1 bash tools/gpu_r2_f.sh
