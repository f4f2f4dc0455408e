"""Quick timing of the decode step (Llama-2-7B Q4, Q8 KV, ctx 2048) under the tuning knobs of the environment:
prints ONE line `tag tok/s ms`. Used by the gpu_r2_*.sh sweeps; bench.py is the contract run."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import inferflow_b200 as ib
tag = sys.argv[1] if len(sys.argv) > 1 else "-"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.zeros(1, device="cuda")
spec = ib.ModelSpec(max_ctx=2048, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(2047, True)
st = torch.cuda.Stream(); sh = st.cuda_stream
with torch.cuda.stream(st):
    for _ in range(5):
        m.decode(1, 2047, stream=sh)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        m.decode(1, 2047, stream=sh)
    e1.record(st); st.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{tag} {1000.0 / ms:.1f} tok/s {ms:.4f} ms/step step_kernel={m.step_status()}", flush=True)
