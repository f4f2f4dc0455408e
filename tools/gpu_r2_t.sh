#!/bin/bash
# boundary proof (reference CUDA engine stock vs with our GEMV/Quantize linked in), tcgen05 GEMM ncu capture,
# BASELINE's 70B shapes on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -x -q > gpurun_out/r2t_boundary.log 2>&1
echo "boundary rc=$?"; tail -15 gpurun_out/r2t_boundary.log
cat > /tmp/ncu_gemm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, orc
import inferflow_b200 as ib
for dt in (orc.Q4_B32T1A, orc.Q3H_B64T1):
    tokens, cy, cx = 4096, 11008, 4096
    w = (torch.randn(cy, cx, device="cuda") * 0.05).to(torch.float16)
    wq = torch.empty(cy * ib.row_bytes(dt, cx), dtype=torch.uint8, device="cuda")
    ib.check(ib.lib.ifb_quantize_rows(dt, wq.data_ptr(), w.data_ptr(), cy, cx, None), "q")
    x = torch.randn(tokens, cx, device="cuda").to(torch.float16)
    y = torch.empty(tokens, cy, dtype=torch.float16, device="cuda")
    for _ in range(3):
        ib.lib.ifb_gemm_q(dt, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, None, None)
    torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_q_kernel -c 6 -f -o gpurun_out/r2t_gemm python /tmp/ncu_gemm.py > gpurun_out/r2t_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2t_ncu.log
timeout 200 python tools/gemm_bench.py > gpurun_out/r2t_gemm_bench.jsonl 2> gpurun_out/r2t_gemm_bench.err
echo "gemm bench rc=$?"; tail -12 gpurun_out/r2t_gemm_bench.jsonl | cut -c1-300
timeout 500 python bench.py --steps 40 --warmup 5 --workload llama2_70b_q4 > gpurun_out/r2t_bench70b_n1.json 2> gpurun_out/r2t_bench70b_n1.err
echo "bench 70B N=1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench70b_n1.json')); print(d['value'], d['ms_per_step'], d['decode_path'], d['verify'], d['roofline']['step'])"
tail -3 gpurun_out/r2t_bench70b_n1.err
