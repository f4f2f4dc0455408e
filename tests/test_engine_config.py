"""The engine-shaped host layer reads the reference's own configuration files: the .ini + model-spec json +
llama2.c weight file that tests/ref_engine.py writes for the REFERENCE engine are resolved by ifb_engine_resolve to
the same model, element types and context length (CPU; the GPU half is tests/test_zzzz_gpu_engine.py)."""
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import orc  # noqa: E402
import ref_engine as R  # noqa: E402


def _tree(d, hp, edit=None):
    w = R.synthetic_llama(hp, 3)
    ini = R.write_tree(d, hp, w, threads=2)
    if edit:
        text = open(ini).read()
        for a, b in edit:
            assert a in text
            text = text.replace(a, b)
        open(ini, "w").write(text)
    return ini


def test_resolves_the_files_written_for_the_reference_engine():
    import inferflow_b200 as ib
    hp = dict(dim=64, hidden=128, layers=2, heads=4, kv_heads=2, vocab=96, seq=32)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, hp)
        path, wdt, kvdt, ctx, order, ret = ib.resolve_config(ini)
        assert os.path.samefile(path, os.path.join(d, "data", "models", "synth", "model.bin"))
        assert (wdt, kvdt, ctx, order, ret) == (orc.F16, orc.F16, 32, 0, 1)
        # the data_root_dir can also come from the caller, as the reference tools pass it (LoadConfig's 4th argument)
        path2 = ib.resolve_config(ini, data_root_dir=os.path.join(d, "data", "inferflow") + "/")[0]
        assert os.path.samefile(path, path2)


def test_element_types_follow_the_reference_rules():
    import inferflow_b200 as ib
    hp = dict(dim=64, hidden=128, layers=1, heads=4, kv_heads=4, vocab=96, seq=32)
    with tempfile.TemporaryDirectory() as d:
        # Q4 weights asked for, but every matrix is below tensor_quant_threshold (2000 x 2000): stays fp16, as in the
        # reference (network_builder.cc:1558); a Q8 KV cache is honoured; max_context_len <= 16 -> the default 4096
        ini = _tree(d, hp, [("device_weight_data_type = F16", "device_weight_data_type = Q4"),
                            ("device_kv_cache_data_type = F16", "device_kv_cache_data_type = Q8"),
                            ("max_context_len = 32", "max_context_len = 8")])
        _, wdt, kvdt, ctx, _, _ = ib.resolve_config(ini)
        assert (wdt, kvdt, ctx) == (orc.F16, orc.Q8_B32T2, 4096)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, hp, [("device_weight_data_type = F16", "device_weight_data_type = Q4\ntensor_quant_threshold = 0")])
        assert ib.resolve_config(ini)[1] == orc.Q4_B32T1A
    # the rule is per tensor (network_builder.cc:1558): a grouped-query model whose wk / wv alone fall below the threshold
    # would get mixed element types -- this runtime holds one type for all layer matrices and says so instead of
    # loading the model under another policy; item names are case-insensitive (sslib ConfigData)
    gqa = dict(dim=64, hidden=128, layers=1, heads=4, kv_heads=1, vocab=96, seq=32)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, gqa, [("device_weight_data_type = F16", "Device_Weight_Data_Type = Q4\nTENSOR_QUANT_THRESHOLD = 2000")])
        with pytest.raises(ib.IfbError, match="per-tensor weight types"):      # wq 4096 >= 2000 > wk 1024 elements
            ib.resolve_config(ini)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, gqa, [("device_weight_data_type = F16", "device_weight_data_type = Q4\ntensor_quant_threshold = 1000")])
        assert ib.resolve_config(ini)[1] == orc.Q4_B32T1A


def test_config_errors_are_reported():
    import inferflow_b200 as ib
    hp = dict(dim=64, hidden=128, layers=1, heads=4, kv_heads=4, vocab=96, seq=32)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, hp, [("device_weight_data_type = F16", "device_weight_data_type = Q9")])
        with pytest.raises(ib.IfbError, match="device_weight_data_type"):
            ib.resolve_config(ini)
    with tempfile.TemporaryDirectory() as d:
        ini = _tree(d, hp)
        spec = os.path.join(d, "data", "models", "synth", "model_spec.json")
        text = open(spec).read()
        open(spec, "w").write(text.replace('"llama2.c"', '"safetensors"'))
        with pytest.raises(ib.IfbError, match="not read by this runtime"):
            ib.resolve_config(ini)
        os.remove(spec)
        with pytest.raises(ib.IfbError, match="model-spec"):
            ib.resolve_config(ini)
    with pytest.raises(ib.IfbError, match="cannot read"):
        ib.resolve_config("/nonexistent/x.ini")
