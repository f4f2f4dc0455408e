"""Prompt throughput of the one-pass prompt (ifb_model_prefill) on Llama-2-7B shapes (synthetic weights): tokens/s for
prompts of 1k and 4k tokens, Q4_B32T1A and Q3H_B64T1 (BASELINE.json configs[2]: 'Q3.5, 4k prefill'), against feeding the
prompt token by token through the decode path (what round 1 did). FLOPs per SURVEY 8d: 2 * 6.476e9 * T + 2 * L * T^2 * d."""
import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import inferflow_b200 as ib
import orc
torch.zeros(1, device="cuda")
for wdt in (orc.Q4_B32T1A, orc.Q3H_B64T1):
    spec = ib.ModelSpec(max_ctx=4096, weight_dtype=wdt, kv_dtype=orc.Q8_B32T2, **ib.LLAMA2_7B)
    m = ib.Model(spec); m.init_synthetic(1234)
    for T in (1024, 4096):
        prompt = [(13 * i + 5) % spec.vocab for i in range(T)]
        m.prefill(prompt, 0); ib.check(ib.lib.ifb_device_synchronize(), "sync")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            m.prefill(prompt, 0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2 * 6.476e9 * T + 2 * 32 * T * T * 4096 / 2 * 2
        print(json.dumps({"weights": orc.NAMES[wdt], "prompt_tokens": T, "ms": ms, "prompt_tokens_per_s": T / ms * 1e3,
                          "TFLOPs": flops / ms / 1e9}), flush=True)
    # token by token (decode path), 64 tokens timed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.kv_reset(0)
    for p in range(8):
        m.decode(5, p)
    torch.cuda.synchronize(); e0.record()
    for p in range(8, 72):
        m.decode(5, p)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"weights": orc.NAMES[wdt], "token_by_token_tokens_per_s": 64 / e0.elapsed_time(e1) * 1e3}), flush=True)
    m.close()
