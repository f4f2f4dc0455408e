"""Batched decode throughput (BASELINE config 5's batch dimension on the dense 7B model): n sequences at context 2047
advance one token per step through ifb_model_decode_batch; tokens/s = n / step time. CUDA events, after warm-up."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import inferflow_b200 as ib
torch.zeros(1, device="cuda")
CTX = 2048
spec = ib.ModelSpec(max_ctx=CTX, kv_slots=8, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(CTX - 1, True)      # slot 0 filled; the other slots hold zeros: same bytes streamed
st = torch.cuda.Stream(); sh = st.cuda_stream
for n in (1, 2, 4, 8):
    toks = np.arange(1, n + 1, dtype=np.int32); slots = np.arange(n, dtype=np.int32); pos = np.full(n, CTX - 1, dtype=np.int32)
    def step():
        ib.check(ib.lib.ifb_model_decode_batch(m._h, toks.ctypes.data, slots.ctypes.data, pos.ctypes.data, n, sh), "batch")
    with torch.cuda.stream(st):
        for _ in range(3):
            step()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        reps = 20
        for _ in range(reps):
            step()
        e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wb = m.weight_bytes; kvb = m.kv_bytes_per_token(CTX) * n
    print(json.dumps({"workload": "Llama-2-7B Q4_B32T1A, Q8 KV, ctx 2048, batched decode", "sequences": n, "ms_per_step": ms,
                      "tokens_per_s": 1000.0 * n / ms, "bytes_per_step": wb + kvb, "achieved_GBps": (wb + kvb) / ms / 1e6}), flush=True)
m.close()
