"""Our CUDA decode runtime against the REFERENCE ENGINE's own outputs (tests/golden/engine_stories110m.npz, made by
tests/golden/make_golden_engine.py from the compiled reference): the stories110M-shaped llama of BASELINE
configs[0] (fp16 weights, fp16 KV cache, llama2.c RoPE order), same seeded weights loaded through
ifb_model_load_tensor."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import orc  # noqa: E402
import ref_engine as R  # noqa: E402
import make_golden_engine as G  # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _read(ib, m, vocab):
    ib.check(ib.lib.ifb_device_synchronize(), "sync")
    logits = np.empty(vocab, dtype=np.float16)
    ib.check(ib.lib.ifb_memcpy_d2h(logits.ctypes.data, m.logits_ptr, vocab * 2), "d2h")
    nxt = np.empty(1, dtype=np.int32)
    ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
    return logits.astype(np.float64), int(nxt[0])


def test_decode_matches_reference_engine_tokens_and_logits():
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    g = np.load(os.path.join(HERE, "golden", "engine_stories110m.npz"))
    hp = G.HP
    spec = ib.ModelSpec(layers=hp["layers"], d_model=hp["dim"], d_ffn=hp["hidden"], heads=hp["heads"],
                        kv_heads=hp["kv_heads"], vocab=hp["vocab"], max_ctx=hp["seq"], weight_dtype=orc.F16,
                        kv_dtype=orc.F16, lm_head_dtype=orc.F16, rope_order=0)
    m = ib.Model(spec)
    for k, v in R.synthetic_llama(hp, int(g["seed"])).items():
        m.load_tensor(k, v.reshape(1, -1) if v.ndim == 1 else v)
    prompt, toks = g["prompt"].tolist(), g["tokens"].tolist()
    seq = prompt + toks
    # 1. teacher-forced along the reference's tokens: logits at the reference's top-16 and at the probe ids
    worst = 0.0
    for i in range(len(seq) - 1):
        m.decode(int(seq[i]), i)
        lg, nxt = _read(ib, m, spec.vocab)
        k = i - (len(prompt) - 1)
        if k < 0:
            continue
        rng_ = np.abs(lg).max()
        for ids, val in ((g["top_ids"][k], g["top_val"][k]), (g["probe_ids"], g["probe_val"][k])):
            err = np.abs(lg[ids] - val).max()
            worst = max(worst, err / rng_)
            assert err <= 2 ** -7 * rng_, f"step {k}: logits differ from the reference engine by {err:.3e} (range {rng_:.2f})"
        if g["top_val"][k][0] - g["top_val"][k][1] > 2 ** -5 * rng_:
            assert nxt == toks[k], f"step {k}: greedy token {nxt} != reference {toks[k]}"
    # 2. free-running on the device (graph replay, device-side argmax): the reference's token stream
    m.kv_reset(0)
    for i, t in enumerate(prompt):
        m.decode(int(t), i)
    out = []
    for _ in range(len(toks)):
        _, nxt = _read(ib, m, spec.vocab)
        out.append(nxt)
        m.decode_next()
    assert out == toks, f"token stream {out} != reference {toks}"
    m.close()
    print(f"worst relative logit error vs the reference engine: {worst:.2e}")
