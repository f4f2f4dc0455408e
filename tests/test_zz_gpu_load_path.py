"""The load path on the GPU (host fp32 -> fp16 -> on-device quantizer -> device tensor) against the oracle's bytes."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("dt", [orc.Q4_B32T1A, orc.Q8_B32T2, orc.Q6_B64T1, orc.Q3H_B64T1, orc.Q2_B32T1A, orc.F16],
                         ids=["Q4_B32T1A", "Q8_B32T2", "Q6_B64T1", "Q3H_B64T1", "Q2_B32T1A", "F16"])
def test_load_tensor_quantizes_like_the_reference(dt):
    """the load path (host fp32 -> fp16 -> on-device quantizer -> device tensor; DeviceTensorBuilder::Build,
    device_tensor_builder.cu:161-330): the bytes on the device are the oracle's (= the reference's) quantization of
    the fp16-rounded matrix, for every tensor of a layer"""
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    spec = ib.ModelSpec(layers=1, d_model=256, d_ffn=512, heads=4, kv_heads=2, vocab=128, max_ctx=64, weight_dtype=dt,
                        lm_head_dtype=orc.F16)
    m = ib.Model(spec)
    rng = np.random.default_rng(5)
    o = orc.oracle()
    hd = spec.d_model // spec.heads
    shapes = {"wq": (spec.heads * hd, 256), "wk": (spec.kv_heads * hd, 256), "wv": (spec.kv_heads * hd, 256),
              "wo": (256, spec.heads * hd), "w1": (512, 256), "w3": (512, 256), "w2": (256, 512)}
    for name, (r, c) in shapes.items():
        w = (0.05 * rng.standard_normal((r, c))).astype(np.float32)
        m.load_tensor(f"layers.0.{name}", w)
        got = m.read_tensor(f"layers.0.{name}")
        w16 = w.astype(np.float16)
        want = w16.reshape(-1).view(np.uint8) if dt == orc.F16 else o.quantize(dt, w16.astype(np.float32)).reshape(-1)
        assert np.array_equal(got[: want.size], want), f"{name}: {np.count_nonzero(got[:want.size] != want)} bytes differ"
    for name, n in (("attention_norm", 256), ("ffn_norm", 256)):
        v = (1 + 0.1 * rng.standard_normal((1, n))).astype(np.float32)
        m.load_tensor(f"layers.0.{name}", v)
        assert np.array_equal(m.read_tensor(f"layers.0.{name}").view(np.float16)[:n], v.astype(np.float16).reshape(-1))
    with pytest.raises(ib.IfbError):
        m.load_tensor("layers.0.wq", np.zeros((3, 256), np.float32))      # wrong shape: reported, not swallowed
    with pytest.raises(ib.IfbError):
        m.load_tensor("layers.7.wq", np.zeros((256, 256), np.float32))    # no such layer
    m.close()
