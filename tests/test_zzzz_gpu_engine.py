"""GPU half of tests/test_engine_config.py: the SAME directory tree (ini + model-spec json + llama2.c weights +
tokenizer) that drives the reference engine drives ifb_engine_* -- and the greedy token stream is the reference
engine's (tests/golden/engine_stories110m.npz), with the logits of every step within the model tolerance."""
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

pytestmark = pytest.mark.gpu


def test_same_config_same_tokens_as_the_reference_engine():
    torch = pytest.importorskip("torch")
    import inferflow_b200 as ib
    torch.zeros(1, device="cuda")
    g = np.load(os.path.join(HERE, "golden", "engine_stories110m.npz"))
    hp = G.HP
    w = R.synthetic_llama(hp, int(g["seed"]))
    with tempfile.TemporaryDirectory() as d:
        ini = R.write_tree(d, hp, w, threads=2)          # exactly what make_golden_engine.py gave the reference
        eng = ib.Engine(ini)
    assert eng.vocab == hp["vocab"]
    toks, logits = eng.generate(g["prompt"].tolist(), len(g["tokens"]), want_logits=True)
    assert toks == g["tokens"].tolist()
    for k in range(len(toks)):
        rng_ = np.abs(logits[k]).max()
        err = np.abs(logits[k][g["top_ids"][k]] - g["top_val"][k]).max()
        assert err <= 2 ** -7 * rng_, f"step {k}: logits differ from the reference engine by {err:.3e}"
    # engine state machine: a second query while one is live is refused (0 = busy), removal frees the slot
    qid = eng.add_query([1, 2, 3])
    assert qid > 0 and eng.add_query([4, 5]) == 0
    eng.remove_query(qid)
    assert eng.add_query([4, 5]) > qid
    eng.close()
