#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_batch.py -x -q -s > gpurun_out/r2v_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2v_tests.log | cut -c1-400
timeout 300 python tools/bench_batch.py > gpurun_out/r2v_batch.jsonl 2> gpurun_out/r2v_batch.err
echo "batch bench rc=$?"; cat gpurun_out/r2v_batch.jsonl | cut -c1-300; tail -3 gpurun_out/r2v_batch.err
timeout 200 python tools/gemm_bench.py 2>/dev/null | grep '"tokens": 8,' > gpurun_out/r2v_gemm8.jsonl
cat gpurun_out/r2v_gemm8.jsonl | cut -c1-260
