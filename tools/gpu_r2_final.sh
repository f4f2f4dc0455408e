#!/bin/bash
# round-2 validation: the whole GPU suite, the contract bench (N=1), BASELINE's Q3.5 configuration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/r2z_tests.log
timeout 400 python bench.py --steps 200 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
echo "bench rc=$?"; head -c 1500 gpurun_out/r2z_bench_n1.json; echo; tail -2 gpurun_out/r2z_bench_n1.err
timeout 300 python bench.py --steps 100 --warmup 5 --workload llama2_7b_q3h > gpurun_out/r2z_bench_q3h.json 2> gpurun_out/r2z_bench_q3h.err
echo "bench q3h rc=$?"; head -c 700 gpurun_out/r2z_bench_q3h.json; echo; tail -2 gpurun_out/r2z_bench_q3h.err
