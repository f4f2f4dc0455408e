#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 120 > gpurun_out/r2l_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -25 gpurun_out/r2l_gemm.log
timeout 600 python tools/gemm_bench.py > gpurun_out/r2l_gemm_bench.jsonl 2> gpurun_out/r2l_gemm_bench.err
echo "gemm bench rc=$?"; cat gpurun_out/r2l_gemm_bench.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['w'],d['tokens'],d['cy'],d['cx'],'ours %.3f ms %.0f TF | ref %.3f ms %.0f TF | x%.2f err %.2e'%(d['ours_ms'],d['ours_TFLOPs'],d['dequant_cublas_transpose_ms'],d['ref_TFLOPs'],d['speedup'],d['max_err']))
"; tail -3 gpurun_out/r2l_gemm_bench.err
