#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py "tests/test_gpu_step.py::test_one_llama2_70b_shaped_layer_matches_oracle" -q > gpurun_out/r2u_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2u_tests.log
timeout 500 python bench.py --steps 40 --warmup 5 --workload llama2_70b_q4 > gpurun_out/r2u_bench70b_n1.json 2> gpurun_out/r2u_bench70b_n1.err
echo "bench 70B N=1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2u_bench70b_n1.json')); print(d['value'], d['ms_per_step'], d['decode_path'], d['verify'], d['roofline']['step'])"
tail -3 gpurun_out/r2u_bench70b_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2u_launches.csv
