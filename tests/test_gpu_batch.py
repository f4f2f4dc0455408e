"""Batched decode (ifb_model_decode_batch): several sequences, each in its own KV slot at its own position, advance by
one token per step -- the step the reference's InferenceEngine::Infer runs over all live queries
(inference_engine.cc:1054-1124, query_state_table.h). Every row of the batch must be what the sequence gets alone:
  * against the prompt pass of the same sequence (the same tensor-core GEMM arithmetic; attention by another kernel,
    split-K partial sums in another order): 2^-8 of the logit range;
  * against token-by-token decode (dp4a GEMV on Q8-quantized activations): the prefill-vs-decode bound, 2^-5."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _single_logits(ib, m, spec):
    ib.check(ib.lib.ifb_device_synchronize(), "sync")
    out = np.empty(spec.vocab, dtype=np.float16)
    ib.check(ib.lib.ifb_memcpy_d2h(out.ctypes.data, m.logits_ptr, spec.vocab * 2), "d2h")
    return out.astype(np.float64)


@pytest.mark.parametrize("cfg", [
    dict(weight_dtype=orc.Q4_B32T1A, kv_dtype=orc.Q8_B32T2, heads=4, kv_heads=4, d_model=512, d_ffn=1024),
    dict(weight_dtype=orc.Q3H_B64T1, kv_dtype=orc.F16, heads=8, kv_heads=2, d_model=512, d_ffn=1408, weight_repack=1),
], ids=["Q4-kvQ8", "Q3H-kvF16-gqa"])
def test_batched_decode_rows_equal_the_sequences_alone(cfg):
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    base = dict(layers=2, vocab=1024, max_ctx=64, **cfg)
    bm = ib.Model(ib.ModelSpec(kv_slots=8, **base))
    one = ib.Model(ib.ModelSpec(**base))
    for m in (bm, one):
        m.init_synthetic(seed=31)
    assert bm.batch_supported and not one.batch_supported
    spec = one.spec
    nseq, steps = 6, 12
    seqs = [[(11 * s + 5 * i + 2) % spec.vocab for i in range(steps)] for s in range(nseq)]
    slots = [5, 0, 7, 2, 3, 6]                   # any distinct slots
    # the sequences alone: prompt pass of the first k+1 tokens (same GEMM arithmetic), and token-by-token decode
    ref_pf = [[None] * steps for _ in range(nseq)]
    ref_dec = [[None] * steps for _ in range(nseq)]
    for s in range(nseq):
        for k in range(steps):
            one.prefill(seqs[s][:k + 1], 0)
            ref_pf[s][k] = _single_logits(ib, one, spec)
        for k in range(steps):
            one.decode(seqs[s][k], k)
            ref_dec[s][k] = _single_logits(ib, one, spec)
    # batched, staggered: sequence s joins at step s, so the rows of a step sit at different positions
    worst_pf = worst_dec = 0.0
    for t in range(steps + nseq - 1):
        rows = [(s, t - s) for s in range(nseq) if 0 <= t - s < steps]
        toks = [seqs[s][k] for s, k in rows]
        logits, nxt = bm.decode_batch(toks, [slots[s] for s, _ in rows], [k for _, k in rows])
        for i, (s, k) in enumerate(rows):
            got = logits[i].astype(np.float64)
            rng_ = np.abs(ref_pf[s][k]).max()
            e1, e2 = np.abs(got - ref_pf[s][k]).max() / rng_, np.abs(got - ref_dec[s][k]).max() / rng_
            worst_pf, worst_dec = max(worst_pf, e1), max(worst_dec, e2)
            assert e1 <= 2 ** -8, f"step {t} sequence {s} position {k}: differs from its prompt pass by {e1:.3e} of the range"
            assert e2 <= 2 ** -5, f"step {t} sequence {s} position {k}: differs from token-by-token decode by {e2:.3e}"
            assert nxt[i] == int(np.argmax(logits[i].astype(np.float32)))
    print(f"worst vs prompt pass {worst_pf:.2e}, vs decode {worst_dec:.2e} of the logit range")
    # argument checks: a slot named twice, more rows than slots, a position past the cache
    with pytest.raises(ib.IfbError):
        bm.decode_batch([1, 2], [3, 3], [0, 0])
    with pytest.raises(ib.IfbError):
        bm.decode_batch([1] * 9, list(range(9)), [0] * 9)
    with pytest.raises(ib.IfbError):
        bm.decode_batch([1], [0], [64])
    with pytest.raises(ib.IfbError):
        one.decode_batch([1], [0], [0])          # no KV slots reserved
    bm.close(); one.close()
