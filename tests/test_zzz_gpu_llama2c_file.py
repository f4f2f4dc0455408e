"""GPU half of tests/test_llama2c_file.py: the golden model written as a llama2.c FILE, loaded by the C++ runtime
(ifb_model_load_llama2c), must reproduce the reference engine's token stream (the engine read the same format)."""
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import ref_engine as R  # noqa: E402
import make_golden_engine as G  # noqa: E402


@pytest.mark.gpu
def test_file_to_runtime_reproduces_the_reference_engine():
    """write the golden model as a llama2.c FILE, load it with ifb_model_load_llama2c, decode: the token stream of the
    reference engine (tests/golden/engine_stories110m.npz), which read the same format"""
    torch = pytest.importorskip("torch")
    import inferflow_b200 as ib
    import orc
    torch.zeros(1, device="cuda")
    g = np.load(os.path.join(HERE, "golden", "engine_stories110m.npz"))
    hp = G.HP
    w = R.synthetic_llama(hp, int(g["seed"]))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.bin")
        R.write_llama2c(path, hp, w)
        m = ib.Model.from_llama2c(path, weight_dtype=orc.F16, kv_dtype=orc.F16)
    assert (m.spec.layers, m.spec.d_model, m.spec.vocab, m.spec.max_ctx) == (hp["layers"], hp["dim"], hp["vocab"], hp["seq"])
    prompt, toks = g["prompt"].tolist(), g["tokens"].tolist()
    for i, t in enumerate(prompt):
        m.decode(int(t), i)
    out = []
    for _ in range(len(toks)):
        ib.check(ib.lib.ifb_device_synchronize(), "sync")
        nxt = np.empty(1, dtype=np.int32)
        ib.check(ib.lib.ifb_memcpy_d2h(nxt.ctypes.data, m.next_token_ptr, 4), "d2h")
        out.append(int(nxt[0]))
        m.decode_next()
    assert out == toks
    m.close()
