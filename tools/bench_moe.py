"""Mixtral-8x7B-shaped mixture-of-experts decode (BASELINE configs[4] shape, batch 1): tokens/s and bytes read per
token. Synthetic weights generated on the device; Q4_B32T1A weights, Q8_B32T2 KV, ctx 2048, greedy, CUDA graph."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import inferflow_b200 as ib  # noqa: E402

CTX = 2048
spec = ib.ModelSpec(layers=32, d_model=4096, d_ffn=14336, heads=32, kv_heads=8, vocab=32000, max_ctx=CTX,
                    weight_dtype=ib.DT.Q4_B32T1A, kv_dtype=ib.DT.Q8_B32T2, lm_head_dtype=ib.DT.F16,
                    experts=8, experts_top_k=2)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
sh = stream.cuda_stream
m = ib.Model(spec)
m.init_synthetic(seed=1234, stream=sh)
m.kv_reset(CTX - 1, fill_synthetic=True, seed=99, stream=sh)
stream.synchronize()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
with torch.cuda.stream(stream):
    for _ in range(5):
        m.decode(1, CTX - 1, stream=sh)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        m.decode(1, CTX - 1, stream=sh)
    e1.record(stream)
    stream.synchronize()
ms = e0.elapsed_time(e1) / steps
nbytes = m.weight_bytes + m.kv_bytes_per_token(CTX)
ids = m.read_tensor("moe_ids", 8).view("int32").tolist()
print(json.dumps({"workload": "Mixtral-8x7B-shaped MoE decode, Q4_B32T1A, Q8 KV, ctx 2048, batch 1, top-2 of 8 experts",
                  "tokens_per_s": 1000.0 / ms, "ms_per_step": ms, "bytes_per_token": nbytes,
                  "achieved_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_6569.6": nbytes / (ms * 1e-3) / 1e9 / 6569.6,
                  "launches_per_step": m.launches_per_step, "last_layer_experts": ids}))
