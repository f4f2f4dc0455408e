#!/bin/bash
# ncu --set full of ONE launch of the step kernel (IFB_STEP_PRIV selects the variant; default: private pipelines)
mkdir -p gpurun_out
cat > /tmp/ncu_step.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import inferflow_b200 as ib
torch.zeros(1, device="cuda")
spec = ib.ModelSpec(max_ctx=2048, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(2047, True)
for _ in range(4):
    m.decode(1, 2047)
ib.check(ib.lib.ifb_device_synchronize(), "sync")
print(m.step_status())
PY
IFB_STEP=1 IFB_STEP_PRIV=${IFB_STEP_PRIV:-1} IFB_NO_GRAPH=1 IFB_STEP_TIMEOUT_MS=60000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 2 -c 1 -f -o gpurun_out/r2q_step python /tmp/ncu_step.py > gpurun_out/r2q_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2q_ncu.log; ls -la gpurun_out/r2q_step.ncu-rep
