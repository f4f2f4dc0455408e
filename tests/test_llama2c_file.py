"""The llama2.c model-file reader of the C++ runtime (ifb_llama2c_probe / ifb_model_load_llama2c): the SAME file the
reference engine loads (tests/ref_engine.py writes it, ModelReader::LoadModel_Llama2DotC reads it). CPU part: header
parsing, size validation, clean failures (the GPU half is tests/test_zzz_gpu_llama2c_file.py)."""
import os
import struct
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import ref_engine as R  # noqa: E402
import make_golden_engine as G  # noqa: E402


def _small(shared):
    hp = dict(dim=64, hidden=128, layers=2, heads=4, kv_heads=2, vocab=96, seq=32)
    w = R.synthetic_llama(hp, 3)
    if shared:
        del w["output"]
    return hp, w


@pytest.mark.parametrize("shared", [False, True], ids=["own_classifier", "shared_classifier"])
def test_probe_reads_what_the_writer_wrote(shared):
    import inferflow_b200 as ib
    hp, w = _small(shared)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.bin")
        R.write_llama2c(path, hp, w)
        hv, sh = ib.probe_llama2c(path)
        assert hv == (hp["dim"], hp["hidden"], hp["layers"], hp["heads"], hp["kv_heads"], hp["vocab"], hp["seq"])
        assert sh == shared
        # a truncated file is refused with a message, not read past its end
        data = open(path, "rb").read()
        open(path, "wb").write(data[: len(data) - 8])
        with pytest.raises(ib.IfbError, match="holds"):
            ib.probe_llama2c(path)


def test_probe_rejects_garbage():
    import inferflow_b200 as ib
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.bin")
        open(p, "wb").write(struct.pack("<7i", 64, 128, 2, 5, 2, 96, 32) + bytes(1 << 20))   # 64 % 5 != 0
        with pytest.raises(ib.IfbError, match="implausible"):
            ib.probe_llama2c(p)
        open(p, "wb").write(b"\x01\x02")
        with pytest.raises(ib.IfbError):
            ib.probe_llama2c(p)
        with pytest.raises(ib.IfbError, match="cannot open"):
            ib.probe_llama2c(os.path.join(d, "missing.bin"))


def test_loading_without_a_gpu_fails_cleanly():
    """no CUDA device here: the loader must come back with an error code and a message, not crash"""
    import inferflow_b200 as ib
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    hp, w = _small(False)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.bin")
        R.write_llama2c(path, hp, w)
        with pytest.raises(ib.IfbError):
            ib.Model.from_llama2c(path)
