"""The tensor-parallel restatements of the model oracle (tests/orc_model.py, `tp=`) on CPU: the shard sums of the
row-parallel products -- dense wo / w2 and, for mixture-of-experts layers, every selected expert's w2 folded per rank --
must be the unsharded model up to the rounding of the extra fp16 stores (the GPU tests pin the kernels to exactly these
restatements: tests/tp_worker.py)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import orc
from orc_model import OracleModel


def _model(experts, tp, wdt=orc.Q4_B32T1A, seed=3):
    s = NS(layers=2, vocab=256, max_ctx=32, weight_dtype=wdt, kv_dtype=orc.Q8_B32T2, lm_head_dtype=orc.F16, heads=4,
           kv_heads=2, d_model=256, d_ffn=512, rope_theta=10000.0, rope_order=2, norm_eps=1e-5, kq_scale=1.0,
           experts=experts, experts_top_k=2 if experts else 0, moe_norm_top_k_prob=1)
    rng = np.random.default_rng(seed)
    o = orc.oracle()
    hd = s.d_model // s.heads

    def q(rows, cols, std=0.05):
        w = (std * rng.standard_normal((rows, cols))).astype(np.float16).astype(np.float32)
        return o.quantize(wdt, w).reshape(-1)

    def h(n, mean=0.0, std=1.0):
        return (mean + std * rng.standard_normal(n)).astype(np.float16).view(np.uint8)
    t = {"tok_embeddings": h(s.vocab * s.d_model), "output_norm": h(s.d_model, 1.0, 0.1), "output": h(s.vocab * s.d_model, 0.0, 0.05)}
    for l in range(s.layers):
        p = f"layers.{l}."
        t[p + "attention_norm"] = h(s.d_model, 1.0, 0.1)
        t[p + "ffn_norm"] = h(s.d_model, 1.0, 0.1)
        t[p + "wq"] = q(s.heads * hd, s.d_model); t[p + "wk"] = q(s.kv_heads * hd, s.d_model); t[p + "wv"] = q(s.kv_heads * hd, s.d_model)
        t[p + "wo"] = q(s.d_model, s.heads * hd)
        if experts:
            t[p + "gate"] = h(experts * s.d_model, 0.0, 0.3)
            for e in range(experts):
                t[p + f"experts.{e}.w1"] = q(s.d_ffn, s.d_model); t[p + f"experts.{e}.w3"] = q(s.d_ffn, s.d_model)
                t[p + f"experts.{e}.w2"] = q(s.d_model, s.d_ffn)
        else:
            t[p + "w1"] = q(s.d_ffn, s.d_model); t[p + "w3"] = q(s.d_ffn, s.d_model); t[p + "w2"] = q(s.d_model, s.d_ffn)
    return OracleModel(s, t, tp=tp)


@pytest.mark.parametrize("experts", [0, 4], ids=["dense", "moe"])
@pytest.mark.parametrize("tp", [2, 4])
def test_shard_sums_are_the_unsharded_model_up_to_fp16_stores(experts, tp):
    a, b = _model(experts, 1), _model(experts, tp)
    tok = 7
    for pos in range(5):
        la, lb = a.step(tok, pos), b.step(tok, pos)
        rng_ = np.abs(la).max()
        err = np.abs(la - lb).max()
        assert 0 < err <= 2 ** -6 * rng_, f"pos {pos}: tp {tp} oracle differs from the unsharded oracle by {err:.3e} (range {rng_:.2f})"
        if experts:
            assert a.last_route == b.last_route or err > 0        # the router sees the same x1 up to that rounding
        tok = int(np.argmax(la))
