#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_gemm.py -x -q --timeout 300 > gpurun_out/r2n_prefill.log 2>&1
echo "prefill tests rc=$?"; tail -25 gpurun_out/r2n_prefill.log
timeout 300 python tools/prefill_bench.py > gpurun_out/r2n_prefill_bench.txt 2>&1
echo "prefill bench rc=$?"; tail -12 gpurun_out/r2n_prefill_bench.txt
