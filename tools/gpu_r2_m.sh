#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ncu_gemm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, orc
import inferflow_b200 as ib
dt = orc.Q4_B32T1A; tokens, cy, cx = 4096, 4096, 4096
w = (torch.randn(cy, cx, device="cuda") * 0.05).to(torch.float16)
wq = torch.empty(cy * ib.row_bytes(dt, cx), dtype=torch.uint8, device="cuda")
ib.check(ib.lib.ifb_quantize_rows(dt, wq.data_ptr(), w.data_ptr(), cy, cx, None), "q")
x = torch.randn(tokens, cx, device="cuda").to(torch.float16)
y = torch.empty(tokens, cy, dtype=torch.float16, device="cuda")
for _ in range(3):
    ib.lib.ifb_gemm_q(dt, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, None, None)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_q_kernel -s 1 -c 1 -f -o gpurun_out/r2m_gemm python /tmp/ncu_gemm.py > gpurun_out/r2m_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2m_ncu.log
