"""Model-level parity on the GPU: the fused CUDA-graph decode runtime (ifb_model_*) against a
host oracle of the same layer schedule on the same (synthetic, read back from the device) weights."""
import ctypes as C

import numpy as np
import pytest

import orc
from orc_model import OracleModel

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NAMES = ["attention_norm", "ffn_norm", "wq", "wk", "wv", "wo", "w1", "w2", "w3"]


def _tensors(m, spec):
    t = {n: m.read_tensor(n) for n in ("tok_embeddings", "output_norm", "output")}
    for l in range(spec.layers):
        for n in NAMES:
            t[f"layers.{l}.{n}"] = m.read_tensor(f"layers.{l}.{n}")
    return t


CONFIGS = [
    dict(weight_dtype=orc.Q4_B32T1A, kv_dtype=orc.Q8_B32T2, heads=8, kv_heads=8, d_model=256),
    dict(weight_dtype=orc.Q3H_B64T1, kv_dtype=orc.F16, heads=8, kv_heads=2, d_model=256),
    dict(weight_dtype=orc.Q8_B32T2, kv_dtype=orc.Q8_B32T2, heads=4, kv_heads=4, d_model=256, rope_order=0),
    dict(weight_dtype=orc.Q2_B32T1B, kv_dtype=orc.Q8_B32T2, heads=2, kv_heads=2, d_model=256),   # fp16-activation path
    dict(weight_dtype=orc.Q6_B64T1, kv_dtype=orc.Q8_B32T2, heads=4, kv_heads=4, d_model=512, lm_head_dtype=orc.Q6_B64T1),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[orc.NAMES[c["weight_dtype"]] for c in CONFIGS])
def test_decode_matches_oracle(cfg):
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    spec = ib.ModelSpec(layers=2, d_ffn=512, vocab=512, max_ctx=128, **cfg)
    m = ib.Model(spec)
    m.init_synthetic(seed=7)
    orcm = OracleModel(spec, _tensors(m, spec))
    tok = 5
    toks_gpu, toks_orc = [], []
    for pos in range(6):
        m.decode(tok, pos)
        ib.check(ib.lib.ifb_device_synchronize(), "sync")
        logits = np.empty(spec.vocab, dtype=np.float16)
        ib.check(ib.lib.ifb_memcpy_d2h(logits.ctypes.data, m.logits_ptr, spec.vocab * 2), "d2h")
        nxt = np.empty(1, dtype=np.int32)
        ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
        want = orcm.step(tok, pos)
        scale = np.abs(want).max()
        err = np.abs(logits.astype(np.float64) - want).max()
        # fp16 ties at every stored intermediate compound over the layers: 2^-7 of the logit range
        assert err <= 2 ** -7 * scale, f"pos {pos}: logits differ by {err:.3e} (range {scale:.3e})"
        assert int(nxt[0]) == int(np.argmax(logits.astype(np.float32))), "device argmax != argmax of device logits"
        srt = np.sort(want)
        if srt[-1] - srt[-2] > 4 * err:          # margin filter
            assert int(nxt[0]) == int(np.argmax(want))
        toks_gpu.append(int(nxt[0])); toks_orc.append(int(np.argmax(want)))
        tok = int(nxt[0])
    # one persistent launch (ifb_step.cu) where the model qualifies (head_dim 64/128, dp4a weight types), else
    # embed; per layer: QKV GEMV, attention (+append), wo, GLU, w2; lm_head, argmax, advance
    assert m.launches_per_step == (1 if m.step_status()[0] else 1 + 5 * spec.layers + 3)
    m.close()


def test_graph_replay_matches_eager_and_decode_next():
    """the CUDA-graph replay (device-side position/token) equals stepping with host-provided inputs"""
    import os
    import inferflow_b200 as ib
    spec = ib.ModelSpec(layers=3, d_model=256, d_ffn=768, heads=4, kv_heads=2, vocab=1024, max_ctx=256)
    outs = []
    for mode in ("graph", "chain"):
        m = ib.Model(spec)
        m.init_synthetic(seed=3)
        toks = []
        if mode == "graph":
            tok = 1
            for pos in range(8):
                m.decode(tok, pos)
                ib.check(ib.lib.ifb_device_synchronize(), "sync")
                nxt = np.empty(1, dtype=np.int32)
                ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
                tok = int(nxt[0]); toks.append(tok)
        else:
            m.decode(1, 0)
            for _ in range(8):
                ib.check(ib.lib.ifb_device_synchronize(), "sync")
                nxt = np.empty(1, dtype=np.int32)
                ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
                toks.append(int(nxt[0]))
                m.decode_next()
        outs.append(toks)
        m.close()
    assert outs[0] == outs[1]


MOE = [dict(weight_dtype=orc.Q4_B32T1A, experts=8, experts_top_k=2), dict(weight_dtype=orc.Q8_B32T2, experts=4, experts_top_k=2),
       dict(weight_dtype=orc.Q4_B32T1A, experts=4, experts_top_k=1, moe_norm_top_k_prob=0)]


@pytest.mark.parametrize("cfg", MOE, ids=[f"{orc.NAMES[c['weight_dtype']]}-e{c['experts']}k{c['experts_top_k']}" for c in MOE])
def test_moe_decode_matches_oracle(cfg):
    """mixture-of-experts FFN (ProcessGpuLayer_Moe): router GEMV, softmax + top-k ON THE DEVICE (expert ids bit-exact
    with the host rule), expert FFNs with device-selected weights, half-precision FMA combine"""
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    spec = ib.ModelSpec(layers=2, d_model=256, d_ffn=512, heads=4, kv_heads=4, vocab=512, max_ctx=64,
                        kv_dtype=orc.Q8_B32T2, **cfg)
    m = ib.Model(spec)
    m.init_synthetic(seed=11)
    t = {n: m.read_tensor(n) for n in ("tok_embeddings", "output_norm", "output")}
    for l in range(spec.layers):
        for n in ("attention_norm", "ffn_norm", "wq", "wk", "wv", "wo", "gate"):
            t[f"layers.{l}.{n}"] = m.read_tensor(f"layers.{l}.{n}")
        for e in range(spec.experts):
            for n in ("w1", "w2", "w3"):
                t[f"layers.{l}.experts.{e}.{n}"] = m.read_tensor(f"layers.{l}.experts.{e}.{n}")
    orcm = OracleModel(spec, t)
    tok = 9
    seen = set()
    for pos in range(6):
        m.decode(tok, pos)
        ib.check(ib.lib.ifb_device_synchronize(), "sync")
        logits = np.empty(spec.vocab, dtype=np.float16)
        ib.check(ib.lib.ifb_memcpy_d2h(logits.ctypes.data, m.logits_ptr, spec.vocab * 2), "d2h")
        nxt = np.empty(1, dtype=np.int32)
        ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
        want = orcm.step(tok, pos)
        # the routing of the LAST layer is still in the device buffers: expert ids bit-exact, weights to the half
        ids = m.read_tensor("moe_ids", 4 * spec.experts_top_k).view(np.int32)
        wts = m.read_tensor("moe_wts", 2 * spec.experts_top_k).view(np.float16)
        route = orcm.last_route
        assert [int(i) for i in ids if i >= 0] == [e for e, _ in route], f"pos {pos}: expert ids {ids} vs oracle {route}"
        for (e, w), got in zip(route, wts):
            assert abs(float(got) - w) <= 2 ** -10 * max(w, 2 ** -4), f"pos {pos}: weight of expert {e}: {got} vs {w}"
        seen.update(e for e, _ in route)
        scale = np.abs(want).max()
        err = np.abs(logits.astype(np.float64) - want).max()
        assert err <= 2 ** -7 * scale, f"pos {pos}: logits differ by {err:.3e} (range {scale:.3e})"
        tok = int(nxt[0])
    assert len(seen) >= 2, "degenerate routing: the test never switched experts"
    assert m.launches_per_step == 1 + (6 + 2 * spec.experts_top_k) * spec.layers + 3
    m.close()

