"""The reference ENGINE (CPU build, run through its public API) as the pin for the model-level oracle.

tests/golden/engine_stories110m.npz was produced by tests/golden/make_golden_engine.py from
oracle/_ref/libifref_cpu.so -- the unmodified reference compiled here -- on a seeded stories110M-shaped llama
(BASELINE configs[0] shape). The host oracle of the decode step (tests/orc_model.py: the GPU worker's layer
schedule with every intermediate stored as half) must reproduce the reference's greedy tokens and logits."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import orc  # noqa: E402
import ref_engine as R  # noqa: E402
from orc_model import OracleModel  # noqa: E402
import make_golden_engine as G  # noqa: E402

GOLD = os.path.join(HERE, "golden", "engine_stories110m.npz")


def oracle_spec(hp):
    from types import SimpleNamespace as NS
    return NS(layers=hp["layers"], vocab=hp["vocab"], max_ctx=hp["seq"], weight_dtype=orc.F16, kv_dtype=orc.F16,
              lm_head_dtype=orc.F16, heads=hp["heads"], kv_heads=hp["kv_heads"], d_model=hp["dim"], d_ffn=hp["hidden"],
              rope_theta=10000.0, rope_order=0, norm_eps=1e-5, kq_scale=1.0)


def test_oracle_model_reproduces_reference_engine():
    g = np.load(GOLD)
    hp = G.HP
    assert [hp[k] for k in ("dim", "hidden", "layers", "heads", "kv_heads", "vocab", "seq")] == g["hp"].tolist()
    w = R.synthetic_llama(hp, int(g["seed"]))
    t = {k: v.astype(np.float16).reshape(-1).view(np.uint8) for k, v in w.items()}
    m = OracleModel(oracle_spec(hp), t)
    prompt, toks = g["prompt"].tolist(), g["tokens"].tolist()
    seq = prompt + toks
    worst = 0.0
    for i in range(len(prompt) + 11):                 # teacher-forced along the reference's own tokens
        lg = m.step(int(seq[i]), i)
        k = i - (len(prompt) - 1)
        if k < 0:
            continue
        rng_ = np.abs(lg).max()
        for ids, val in ((g["top_ids"][k], g["top_val"][k]), (g["probe_ids"], g["probe_val"][k])):
            err = np.abs(lg[ids] - val).max()
            worst = max(worst, err / rng_)
            # the CPU engine keeps fp32 activations, the GPU schedule (and this oracle) stores halves: 2^-7 of range
            assert err <= 2 ** -7 * rng_, f"step {k}: logits differ from the reference engine by {err:.3e} (range {rng_:.2f})"
        margin = g["top_val"][k][0] - g["top_val"][k][1]
        if margin > 2 ** -5 * rng_:
            assert int(np.argmax(lg)) == toks[k], f"step {k}: greedy token {int(np.argmax(lg))} != reference {toks[k]}"
    print(f"worst relative logit error vs the reference engine: {worst:.2e}")


@pytest.mark.skipif(not os.path.exists(orc.ref_cpu_path()), reason="oracle/_ref/libifref_cpu.so not built (needs /root/reference)")
def test_golden_is_what_the_reference_engine_produces_now():
    """regenerate with the compiled reference and compare with the committed fixture"""
    g = np.load(GOLD)
    toks, logits, _ = G.run_reference(threads=4)
    assert toks.tolist() == g["tokens"].tolist()
    top = np.take_along_axis(logits, g["top_ids"].astype(np.int64), axis=1)
    assert np.abs(top - g["top_val"]).max() <= 1e-3       # thread count changes ggml's summation order slightly
