"""Prompt processing (rows > 1): the prefill attention kernel against an fp32 restatement of the attention on the cache's
own (dequantized) rows, and the one-pass prompt (ifb_model_prefill: tensor-core GEMMs + prefill attention) against feeding
the same prompt token by token through the decode path.
The two model paths are different arithmetic BY THE REFERENCE'S OWN DESIGN: rows > 1 multiply fp16 activations with the
dequantized weights (Dequantize + cuBLAS, inference_worker.cc:2364-2432), a single row quantizes the activation to
Q8_B32T2 first (Gemv_AX8). So the cross-check is a tolerance check (2^-5 of the logit range after 2 layers, greedy tokens
equal where the margin allows); the exact checks are the operator tests (test_gpu_gemm.py, the attention test below)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CASES = [  # T, pos0, heads, kv_heads, head_dim, kv_dtype
    (2, 0, 4, 4, 128, orc.Q8_B32T2), (17, 5, 4, 2, 128, orc.Q8_B32T2), (128, 0, 8, 8, 64, orc.F16),
    (200, 37, 8, 1, 64, orc.Q8_B32T2), (1024, 0, 2, 2, 128, orc.Q8_B32T2), (65, 63, 4, 4, 128, orc.F16),
]


@pytest.mark.parametrize("case", CASES, ids=[f"T{c[0]}_p{c[1]}_h{c[2]}g{c[2] // c[3]}_hd{c[4]}_{'q8' if c[5] == orc.Q8_B32T2 else 'f16'}" for c in CASES])
def test_attn_prefill_matches_fp32_oracle(case):
    import inferflow_b200 as ib
    T, pos0, heads, kvh, hd, kvdt = case
    ctx = pos0 + T
    max_ctx = (ctx + 63) // 64 * 64 + 64
    kvd = kvh * hd
    g = torch.Generator(device="cuda"); g.manual_seed(T * 7 + pos0)
    k = torch.randn(ctx, kvd, device="cuda", generator=g).to(torch.float16)
    v = torch.randn(ctx, kvd, device="cuda", generator=g).to(torch.float16)
    q = torch.randn(T, heads * hd, device="cuda", generator=g).to(torch.float16)
    nbytes = ib.lib.ifb_kv_cache_bytes(kvdt, kvh, hd, max_ctx)
    kc = torch.zeros(nbytes, dtype=torch.uint8, device="cuda"); vc = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    ib.check(ib.lib.ifb_kv_append(kvdt, kc.data_ptr(), vc.data_ptr(), k.data_ptr(), v.data_ptr(), kvh, hd, max_ctx, 0, ctx, None), "append")
    out = torch.full((T, heads * hd), float("nan"), dtype=torch.float16, device="cuda")
    ib.check(ib.lib.ifb_attn_prefill(kvdt, out.data_ptr(), q.data_ptr(), kc.data_ptr(), vc.data_ptr(), T, pos0, heads, kvh, hd, max_ctx, None), "prefill")
    # the cache's own rows, dequantized as GetKRows / GetVRows return them
    kd = torch.empty(ctx, kvd, dtype=torch.float16, device="cuda"); vd = torch.empty(ctx, kvd, dtype=torch.float16, device="cuda")
    ib.check(ib.lib.ifb_kv_get_rows(kvdt, kd.data_ptr(), kc.data_ptr(), kvh, hd, max_ctx, 0, ctx, None), "get")
    ib.check(ib.lib.ifb_kv_get_rows(kvdt, vd.data_ptr(), vc.data_ptr(), kvh, hd, max_ctx, 0, ctx, None), "get")
    torch.cuda.synchronize()
    G = heads // kvh
    qf = q.float().view(T, heads, hd).permute(1, 0, 2)                        # [h][T][hd]
    kf = kd.float().view(ctx, kvh, hd).permute(1, 0, 2).repeat_interleave(G, 0)
    vf = vd.float().view(ctx, kvh, hd).permute(1, 0, 2).repeat_interleave(G, 0)
    s = qf @ kf.transpose(1, 2) / np.sqrt(hd)
    pos_q = torch.arange(T, device="cuda")[:, None] + pos0
    s = s.masked_fill(torch.arange(ctx, device="cuda")[None, :] > pos_q, float("-inf"))
    want = (torch.softmax(s, -1) @ vf).permute(1, 0, 2).reshape(T, heads * hd)
    err = (out.float() - want).abs().max().item()
    # probabilities enter the tensor core as fp16 (2^-11 relative each), the output is stored as half
    assert np.isfinite(err) and err <= 2 ** -8 * max(1.0, want.abs().max().item()), f"{case}: |out - oracle| = {err:.3e}"


def test_model_prefill_against_token_by_token():
    import inferflow_b200 as ib
    spec = ib.ModelSpec(layers=2, d_model=512, d_ffn=1024, heads=4, kv_heads=2, vocab=1024, max_ctx=256,
                        weight_dtype=orc.Q4_B32T1A, kv_dtype=orc.Q8_B32T2)
    prompt = [(7 * i + 3) % spec.vocab for i in range(70)]

    def read(m):
        ib.check(ib.lib.ifb_device_synchronize(), "sync")
        logits = np.empty(spec.vocab, dtype=np.float16)
        ib.check(ib.lib.ifb_memcpy_d2h(logits.ctypes.data, m.logits_ptr, spec.vocab * 2), "d2h")
        nxt = np.empty(1, dtype=np.int32)
        ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
        return logits.astype(np.float64), int(nxt[0])

    a = ib.Model(spec); a.init_synthetic(seed=21)
    b = ib.Model(spec); b.init_synthetic(seed=21)
    assert a.prefill_supported
    a.prefill(prompt, 0)
    la, ta = read(a)
    for pos, tok in enumerate(prompt):
        b.decode(tok, pos)
    lb, tb = read(b)
    rng_ = np.abs(lb).max()
    err = np.abs(la - lb).max()
    assert err <= 2 ** -5 * rng_, f"prefill vs token-by-token logits differ by {err:.3e} (range {rng_:.3e})"
    srt = np.sort(lb)
    if srt[-1] - srt[-2] > 4 * err:
        assert ta == tb
    # generation continues from the prefilled cache with the decode path
    cont_a = []
    for _ in range(4):
        a.decode_next()
        _, t = read(a)
        cont_a.append(t)
    assert all(0 <= t < spec.vocab for t in cont_a)
    # a second prefill chunk appended behind the first: positions 70..89
    a2 = ib.Model(spec); a2.init_synthetic(seed=21)
    a2.prefill(prompt[:50], 0)
    a2.prefill(prompt[50:], 50)
    l2, _ = read(a2)
    assert np.abs(l2 - la).max() <= 2 ** -9 * rng_, "chunked prefill differs from the one-pass prefill"
    with pytest.raises(ib.IfbError):
        a.prefill([1] * 300, 0)          # beyond max_context_len
    for m in (a, b, a2):
        m.close()
