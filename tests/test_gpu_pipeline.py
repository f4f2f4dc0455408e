"""Layer parallelism (BY_LAYER device groups, NetworkBuilder::SplitGpuLayers, network_builder.cc:2094-2145;
InferenceEngine::Infer_LayerParallelism, inference_engine.cc:1185-1220) and the lossless 3.5-bit -> nibble repacking.
  * two pipeline stages of a model (here on ONE GPU; across GPUs only the copy differs) produce, token for token and
    bit for bit, the logits of the unsplit model;
  * the engine with `devices = 0;0` (two groups of one device) emits the tokens of `devices = 0`;
  * a Q3H_B64T1 model stored as Q4_B64T1 blocks (weight_repack) computes the same bits as the one stored as given."""
import os
import sys
import tempfile

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _logits(ib, m, spec):
    ib.check(ib.lib.ifb_device_synchronize(), "sync")
    out = np.empty(spec.vocab, dtype=np.float16)
    ib.check(ib.lib.ifb_memcpy_d2h(out.ctypes.data, m.logits_ptr, spec.vocab * 2), "d2h")
    nxt = np.empty(1, dtype=np.int32)
    ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
    return out, int(nxt[0])


@pytest.mark.parametrize("cfg", [
    dict(weight_dtype=orc.Q4_B32T1A, kv_dtype=orc.Q8_B32T2, heads=4, kv_heads=4, d_model=512, d_ffn=1024),
    dict(weight_dtype=orc.Q8_B32T2, kv_dtype=orc.F16, heads=8, kv_heads=2, d_model=256, d_ffn=512),
], ids=["Q4-kvQ8", "Q8-kvF16-gqa"])
def test_two_pipeline_stages_equal_the_whole_model(cfg):
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    L = 5                                        # SplitGpuLayers over 2 groups: [0, 3) and [3, 5)
    whole = ib.Model(ib.ModelSpec(layers=L, vocab=512, max_ctx=96, **cfg))
    a = ib.Model(ib.ModelSpec(layers=L, vocab=512, max_ctx=96, layer_begin=0, layer_end=3, **cfg))
    b = ib.Model(ib.ModelSpec(layers=L, vocab=512, max_ctx=96, layer_begin=3, layer_end=5, **cfg))
    for m in (whole, a, b):
        m.init_synthetic(seed=21)
        m.kv_reset(40, fill_synthetic=True, seed=4)
    # the synthetic tensors are a function of the GLOBAL layer: stage b holds layers 3, 4 of the same model
    assert np.array_equal(whole.read_tensor("layers.3.wq"), b.read_tensor("layers.3.wq"))
    assert np.array_equal(whole.read_tensor("layers.4.kcache"), b.read_tensor("layers.4.kcache"))
    spec = whole.spec
    tok = 3
    for pos in range(40, 56):
        whole.decode(tok, pos)
        want, want_tok = _logits(ib, whole, spec)
        a.decode(tok, pos)
        ib.check(ib.lib.ifb_device_synchronize(), "sync")
        ib.check(ib.lib.ifb_memcpy_d2d(b.hidden_in_ptr, a.hidden_out_ptr, spec.d_model * 2), "hand-off")
        b.decode(tok, pos)
        got, got_tok = _logits(ib, b, spec)
        assert np.array_equal(got, want), f"pos {pos}: staged logits differ from the whole model"
        assert got_tok == want_tok
        tok = want_tok
    for m in (whole, a, b):
        m.close()


def test_engine_layer_parallel_groups_emit_the_same_tokens():
    import inferflow_b200 as ib
    import ref_engine as R
    import make_golden_engine as G
    torch.zeros(1, device="cuda")
    g = np.load(os.path.join(HERE, "golden", "engine_stories110m.npz"))
    hp = G.HP
    w = R.synthetic_llama(hp, int(g["seed"]))
    with tempfile.TemporaryDirectory() as d:
        ini = R.write_tree(d, hp, w, threads=2)
        one = ib.Engine(ini)
        text = open(ini).read().replace("devices = 0\n", "Devices = 0;0;0\n")      # item names are case-insensitive
        ini2 = os.path.join(os.path.dirname(ini), "layer_parallel.ini")
        open(ini2, "w").write(text)
        three = ib.Engine(ini2)
        bad = os.path.join(os.path.dirname(ini), "tp.ini")
        open(bad, "w").write(text.replace("Devices = 0;0;0", "devices = 0&0"))
        with pytest.raises(ib.IfbError):
            ib.Engine(bad)                     # tensor-parallel groups: one process per GPU (ifb_rank_plan), said loudly
    prompt = g["prompt"].tolist()
    t1, l1 = one.generate(prompt, 12, want_logits=True)
    t3, l3 = three.generate(prompt, 12, want_logits=True)
    assert t1 == t3 == g["tokens"].tolist()[:12]
    for x, y in zip(l1, l3):
        # the one-stage engine runs the prompt as one prefill pass (tensor-core GEMMs), the staged one token by token:
        # the same model within the model tolerance
        assert np.abs(x - y).max() <= 2 ** -7 * np.abs(x).max()
    one.close(); three.close()


def test_q3h_repacked_as_nibbles_computes_the_same_bits():
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    cfg = dict(layers=2, vocab=512, max_ctx=64, weight_dtype=orc.Q3H_B64T1, kv_dtype=orc.Q8_B32T2, heads=4, kv_heads=2,
               d_model=512, d_ffn=1024, lm_head_dtype=orc.Q3H_B64T1)
    o = orc.oracle()
    for step in ("0", "1"):
        old = os.environ.get("IFB_STEP")
        os.environ["IFB_STEP"] = step
        try:
            plain = ib.Model(ib.ModelSpec(**cfg))
            packed = ib.Model(ib.ModelSpec(weight_repack=1, **cfg))
        finally:
            if old is None:
                del os.environ["IFB_STEP"]
            else:
                os.environ["IFB_STEP"] = old
        assert plain.storage_weight_dtype == orc.Q3H_B64T1 and packed.storage_weight_dtype == orc.Q4_B64T1
        for m in (plain, packed):
            m.init_synthetic(seed=8)
        # the repacked blocks hold the same values: dequantize both forms on the host
        for name, rows, cols in (("layers.1.wq", 512, 512), ("layers.0.w2", 512, 1024), ("output", 512, 512)):
            a = o.dequantize(orc.Q3H_B64T1, plain.read_tensor(name)[: rows * cols // 64 * 32].reshape(rows, -1), cols)
            b = o.dequantize(orc.Q4_B64T1, packed.read_tensor(name)[: rows * cols // 64 * 36].reshape(rows, -1), cols)
            assert np.array_equal(a, b), name
        assert packed.weight_bytes * 32 == plain.weight_bytes * 36       # 36 instead of 32 bytes per 64 weights
        tok = 7
        for pos in range(20):
            plain.decode(tok, pos)
            want, want_tok = _logits(ib, plain, plain.spec)
            packed.decode(tok, pos)
            got, got_tok = _logits(ib, packed, packed.spec)
            assert np.array_equal(got, want) and got_tok == want_tok, f"IFB_STEP={step} pos {pos}"
            tok = want_tok
        # a loaded matrix takes the same route: quantized to Q3H as the reference does, then repacked
        w = (0.05 * np.random.default_rng(3).standard_normal((512, 512))).astype(np.float32)
        for m in (plain, packed):
            m.load_tensor("layers.0.wq", w)
        a = o.dequantize(orc.Q3H_B64T1, plain.read_tensor("layers.0.wq")[: 512 * 8 * 32].reshape(512, -1), 512)
        b = o.dequantize(orc.Q4_B64T1, packed.read_tensor("layers.0.wq")[: 512 * 8 * 36].reshape(512, -1), 512)
        assert np.array_equal(a, b)
        plain.close(); packed.close()
