#!/bin/bash
# round 2, call A: first contact of the persistent step kernel with the hardware
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_step.py -x -q --timeout 300 > gpurun_out/r2a_step.log 2>&1
echo "step tests rc=$?" >> gpurun_out/r2a_step.log
tail -30 gpurun_out/r2a_step.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_step.json 2> gpurun_out/r2a_bench_step.err
echo "bench step rc=$?"; cat gpurun_out/r2a_bench_step.json; tail -5 gpurun_out/r2a_bench_step.err
IFB_STEP=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_multi.json 2> gpurun_out/r2a_bench_multi.err
echo "bench multi rc=$?"; cat gpurun_out/r2a_bench_multi.json
IFB_TRACE=1 timeout 300 python tools/trace_step2.py > gpurun_out/r2a_trace.txt 2>&1
tail -60 gpurun_out/r2a_trace.txt
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 --deselect tests/test_gpu_step.py > gpurun_out/r2a_all.log 2>&1
echo "all gpu tests rc=$?"; tail -15 gpurun_out/r2a_all.log
