#!/bin/bash
# pipeline stages, layer-parallel engine, 3.5-bit repacking; then BASELINE's Q3.5 decode with the repacked weights
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_zzzz_gpu_engine.py tests/test_zz_gpu_load_path.py tests/test_gpu_model.py tests/test_zzz_gpu_llama2c_file.py -x -q > gpurun_out/r2s_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2s_tests.log
true

