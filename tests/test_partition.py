"""Host logic of the partitioner (csrc/ifb_partition.cu) against (1) fixtures traced by hand through the reference's code,
(2) fixtures PRODUCED BY THE COMPILED REFERENCE (tests/golden/partition.json) and (3), where oracle/_ref is present, the
compiled reference itself on 300 random device strings (through its own .ini reader) and 300 random layer splits
(no GPU; nothing under /root/reference is read at run time):
  * InferenceEngine::LoadDeviceGroups, inference_engine.cc:1738-1783 -- the `devices` grammar;
  * the strategy rule, inference_engine.cc:1509-1515;
  * NetworkBuilder::SplitGpuLayers, network_builder.cc:2094-2117 -- layer ranges of the groups;
  * Build{Attn,Ffn,Moe}TensorIdMap + AddLayerTasks_TensorParallel, network_builder.cc:1373-1413, 1594-1688 --
    BY_COL / BY_ROW / DUP per tensor and the per-tensor tensor_quant_threshold rule."""
import ctypes as C
import json
import os
import random
import sys
import tempfile

import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))

BY_LAYER, BY_TENSOR, HYBRID = 0, 1, 2
DUP, BY_ROW, BY_COL = 0, 1, 2


def _lib():
    import inferflow_b200 as ib
    return ib


def _parse(spec):
    ib = _lib()
    ids = (C.c_int32 * 64)()
    ng, gs, st = C.c_int32(), C.c_int32(), C.c_int32()
    rc = ib.lib.ifb_device_groups_parse(spec.encode(), ids, 64, C.byref(ng), C.byref(gs), C.byref(st))
    if rc != 0:
        return rc, ib.lib.ifb_last_error().decode()
    groups = [list(ids[g * gs.value:(g + 1) * gs.value]) for g in range(ng.value)]
    return groups, st.value


# LoadDeviceGroups: split on ",;", trim, skip empty tokens; inside a token split on "&|", trim, skip empty, atoi
GRAMMAR = [
    ("0", [[0]], BY_LAYER),                                   # bin/inferflow_service.ini default: one device
    ("0;1", [[0], [1]], BY_LAYER),                            # group_size 1 -> BY_LAYER whatever the group count
    ("0,1,2,3", [[0], [1], [2], [3]], BY_LAYER),              # ',' separates groups like ';'
    ("0&1", [[0, 1]], BY_TENSOR),                             # one group of several devices
    ("0|1|2|3", [[0, 1, 2, 3]], BY_TENSOR),                   # '|' joins like '&'
    ("0&1&2&3&4&5&6&7", [[0, 1, 2, 3, 4, 5, 6, 7]], BY_TENSOR),
    ("0&1;2&3", [[0, 1], [2, 3]], HYBRID),
    ("0&1&2&3;4&5&6&7", [[0, 1, 2, 3], [4, 5, 6, 7]], HYBRID),
    (" 0 & 1 ;\t2 | 3 ; ", [[0, 1], [2, 3]], HYBRID),          # String::Trim on every piece, trailing separator ignored
    ("0;;1", [[0], [1]], BY_LAYER),                           # empty tokens are skipped
    ("3&1;0&2", [[3, 1], [0, 2]], HYBRID),                    # order kept as written
    ("gpu1&2", [[0, 2]], BY_TENSOR),                          # atoi("gpu1") == 0, as in the reference
]


@pytest.mark.parametrize("spec,groups,strategy", GRAMMAR, ids=[g[0].strip() or "blank" for g in GRAMMAR])
def test_devices_grammar(spec, groups, strategy):
    assert _parse(spec) == (groups, strategy)


def test_devices_grammar_rejects_ragged_groups():
    rc, msg = _parse("0&1;2")
    assert rc != 0 and msg == "All device groups should have the same size: 2 vs. 1"      # the reference's message
    assert _parse("") == ([], BY_LAYER)                       # no device named: no groups (the engine then runs on the CPU)


def _split(a, b, n):
    ib = _lib()
    r, g, k = (C.c_int32 * (2 * n))(), (C.c_int32 * n)(), C.c_int32()
    assert ib.lib.ifb_split_layers(a, b, n, r, g, C.byref(k)) == 0
    return [(r[2 * i], r[2 * i + 1], g[i]) for i in range(k.value)]


def test_split_gpu_layers():
    # layer_num_per_group = ceil(L / groups); start = g * per; end = (g + 1) * per, the last group ends at L
    assert _split(0, 32, 1) == [(0, 32, 0)]
    assert _split(0, 32, 2) == [(0, 16, 0), (16, 32, 1)]
    assert _split(0, 32, 3) == [(0, 11, 0), (11, 22, 1), (22, 32, 2)]
    assert _split(0, 80, 8) == [(i * 10, i * 10 + 10, i) for i in range(8)]
    assert _split(0, 12, 5) == [(0, 3, 0), (3, 6, 1), (6, 9, 2), (9, 12, 3)]          # the 5th group has no layer: dropped
    # the reference's arithmetic as written: a middle group can be handed an end past the last layer (5 over 4:
    # per = 2, group 2 -> [4, 6)); its workers stop at end_layer (inference_engine.cc:1203). ifb_rank_plan clamps.
    assert _split(0, 5, 4) == [(0, 2, 0), (2, 4, 1), (4, 6, 2)]
    assert _split(0, 0, 2) == []
    assert _split(0, 32, 0) == []


def test_rank_plan_hybrid_and_layer_parallel():
    ib = _lib()
    pl = (C.c_int32 * 9)()
    want = {0: [0, 0, 2, 0, 16, -1, 2, 0, HYBRID], 1: [0, 1, 2, 0, 16, -1, 3, 1, HYBRID],
            2: [1, 0, 2, 16, 32, 0, -1, 2, HYBRID], 3: [1, 1, 2, 16, 32, 1, -1, 3, HYBRID]}
    for r, w in want.items():       # group, tp_rank, tp_size, layer_begin, layer_end, prev_rank, next_rank, device, strategy
        assert ib.lib.ifb_rank_plan(b"0&1;2&3", r, 32, pl) == 0 and list(pl) == w
    assert ib.lib.ifb_rank_plan(b"0&1&2&3&4&5&6&7", 5, 80, pl) == 0 and list(pl) == [0, 5, 8, 0, 80, -1, -1, 5, BY_TENSOR]
    assert ib.lib.ifb_rank_plan(b"4;5;6", 1, 32, pl) == 0 and list(pl) == [1, 0, 1, 11, 22, 0, 2, 5, BY_LAYER]
    # 5 layers over 4 groups: group 2 is clamped to [4, 5), group 3 holds nothing and is nobody's neighbour
    assert ib.lib.ifb_rank_plan(b"0;1;2;3", 2, 5, pl) == 0 and list(pl)[:7] == [2, 0, 1, 4, 5, 1, -1]
    assert ib.lib.ifb_rank_plan(b"0;1;2;3", 3, 5, pl) == 0 and list(pl)[3:7] == [0, 0, -1, -1]
    assert ib.lib.ifb_rank_plan(b"0&1", 2, 32, pl) != 0          # rank outside the device list


def test_tensor_partition_types():
    ib = _lib()
    t = lambda n: ib.lib.ifb_tensor_partition_type(n.encode())
    for n in ("wq", "wk", "wv", "qkv", "w1", "w3", "layers.7.wq", "layers.0.W3"):
        assert t(n) == BY_COL            # tensor type 2: split along the output rows
    for n in ("wo", "w2", "layers.31.wo"):
        assert t(n) == BY_ROW            # types 3 and 5: split along the contraction dimension
    for n in ("wq_b", "wk_b", "wv_b", "qkv_b", "w1_b", "w3_b"):
        assert t(n) == BY_COL            # type 1: the biases follow their split outputs
    for n in ("attention_norm", "ffn_norm", "wo_b", "w2_b", "layers.3.moe_gate"):
        assert t(n) == DUP               # type 0
    assert t("tok_embeddings") == -1 and t("output") == -1


def test_tensor_quant_threshold_rule():
    ib = _lib()
    f = ib.lib.ifb_tensor_device_dtype
    thr = 2000 * 2000                    # ModelSpec::tensor_quant_threshold default (model.h:137)
    assert f(4096 * 4096, 1, orc.Q4_B32T1A, thr) == orc.Q4_B32T1A
    assert f(4096 * 4096, 8, orc.Q4_B32T1A, thr) == orc.F16            # 2,097,152 elements per device < 4,000,000
    assert f(11008 * 4096, 8, orc.Q4_B32T1A, thr) == orc.Q4_B32T1A
    assert f(288 * 288, 1, orc.Q8_B32T2, thr) == orc.F16               # the reference's tiny models stay fp16
    assert f(288 * 288, 1, orc.Q8_B32T2, 0) == orc.Q8_B32T2            # tensor_quant_threshold = 0: always quantize
    assert f(100, 1, orc.F16, thr) == orc.F16


# ---- against the COMPILED reference: committed fixtures (tests/golden/partition.json, made by
# tests/golden/make_golden_partition.py from oracle/_ref/libifref_cpu.so) and, where that library is present, live
def _ours_groups(text):
    r = _parse(text)
    if isinstance(r[0], int):
        return {"error": -1}
    return {"groups": r[0]}


def test_grammar_and_layer_split_match_the_compiled_reference_fixtures():
    gold = json.load(open(os.path.join(HERE, "golden", "partition.json")))
    assert len(gold["grammar"]) >= 30 and len(gold["splits"]) >= 16
    for case in gold["grammar"]:
        want = {k: case[k] for k in ("groups", "error") if k in case}
        assert _ours_groups(case["devices"]) == want, repr(case["devices"])
    for case in gold["splits"]:
        got = [[a, b, g] for a, b, g in _split(case["start"], case["end"], case["groups"])]
        assert got == case["ranges"], case


@pytest.mark.skipif(not os.path.exists(orc.ref_cpu_path()), reason="oracle/_ref/libifref_cpu.so not built (needs /root/reference)")
def test_grammar_and_layer_split_match_the_compiled_reference_live():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_partition as G
    lib = G.ref_lib()
    rnd = random.Random(11)
    alphabet = "0123456789  &|;,;&x-"
    with tempfile.TemporaryDirectory() as d:
        for _ in range(300):
            text = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 14))).strip()
            if "=" in text or "\n" in text or text.startswith(("#", ";", "[")):
                continue                     # not an item VALUE of an .ini line any more
            assert _ours_groups(text) == G.ref_groups(lib, text, d), repr(text)
    for _ in range(300):
        a, b, n = 0, rnd.randint(0, 130), rnd.randint(0, 9)
        assert [[x, y, g] for x, y, g in _split(a, b, n)] == G.ref_split(lib, a, b, n), (a, b, n)


@pytest.mark.skipif(not os.path.exists(orc.ref_cpu_path()), reason="oracle/_ref/libifref_cpu.so not built (needs /root/reference)")
def test_ini_names_are_case_insensitive_in_the_reference_as_in_the_engine():
    """sslib ConfigData finds [Transformer_Engine] / DEVICES under the names transformer_engine / devices: the engine's
    .ini reader (csrc/ifb_engine.cu) lower-cases section and item names for the same reason"""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_partition as G
    lib = G.ref_lib()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.ini")
        for text, want in (("[transformer_engine]\nDEVICES = 0;1\n", [[0], [1]]), ("[Transformer_Engine]\nDevices = 0&1\n", [[0, 1]])):
            open(path, "w").write(text)
            ids, sizes = (C.c_int32 * 16)(), (C.c_int32 * 8)()
            n = lib.ifref_load_device_groups(path.encode(), b"transformer_engine", ids, 16, sizes, 8)
            got, k = [], 0
            for g in range(n):
                got.append([ids[k + i] for i in range(sizes[g])])
                k += sizes[g]
            assert got == want
