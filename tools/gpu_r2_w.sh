#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ncu_batch.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import inferflow_b200 as ib
torch.zeros(1, device="cuda")
spec = ib.ModelSpec(max_ctx=2048, kv_slots=8, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(2047, True)
n = 8
toks = np.arange(1, n + 1, dtype=np.int32); slots = np.arange(n, dtype=np.int32); pos = np.full(n, 2047, dtype=np.int32)
for _ in range(2):
    ib.check(ib.lib.ifb_model_decode_batch(m._h, toks.ctypes.data, slots.ctypes.data, pos.ctypes.data, n, None), "batch")
ib.check(ib.lib.ifb_device_synchronize(), "sync")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_batch_launches.csv python /tmp/ncu_batch.py > gpurun_out/r2w_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2w_batch_launches.csv
python tools/summarize_ncu.py launches gpurun_out/r2w_batch_launches.csv | head -30
timeout 300 python tools/bench_batch.py > gpurun_out/r2w_batch.jsonl 2> gpurun_out/r2w_batch.err
cat gpurun_out/r2w_batch.jsonl | cut -c90-250
