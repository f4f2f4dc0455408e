"""The oracle (oracle/oracle.c) against the fixtures generated FROM THE REFERENCE
(tests/golden/make_golden.py) and, when the reference build is present in this
container, against the reference live on fresh random inputs. Bit-exact everywhere:
block bytes, GetInt4 words, expert ids; dequantized floats identical."""
import os

import numpy as np
import pytest

import orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(dt):
    return np.load(os.path.join(GOLD, f"codec_{orc.NAMES[dt]}.npz"))


@pytest.mark.parametrize("dt", orc.ALL_QUANT, ids=[orc.NAMES[d] for d in orc.ALL_QUANT])
def test_quantize_matches_reference_fixture(dt):
    g = _load(dt)
    o = orc.oracle()
    got = o.quantize(dt, g["src"], host_q8=(dt == orc.Q8_B32T2))
    assert got.shape == g["blocks"].shape
    assert np.array_equal(got, g["blocks"]), f"{orc.NAMES[dt]}: quantized bytes differ from the reference"


@pytest.mark.parametrize("dt", orc.ALL_QUANT, ids=[orc.NAMES[d] for d in orc.ALL_QUANT])
def test_dequantize_matches_reference_fixture(dt):
    g = _load(dt)
    got = orc.oracle().dequantize(dt, g["blocks"], g["src"].shape[1])
    assert np.array_equal(got, g["deq"])


@pytest.mark.parametrize("dt", orc.AX8_TYPES, ids=[orc.NAMES[d] for d in orc.AX8_TYPES])
def test_codes_match_reference_getint4(dt):
    g = _load(dt)
    codes, _, _ = orc.oracle().block_codes(dt, g["blocks"], g["src"].shape[1])
    words = g["int4"].reshape(-1).view(np.uint8).reshape(-1, 4)     # element i in byte i-start
    ref_codes = words.reshape(-1, orc.CAP[dt])
    if dt == orc.Q8_B32T2:
        ref_codes = ref_codes.view(np.int8)
    assert np.array_equal(codes, ref_codes.astype(np.int32))


def test_block_sizes():
    o = orc.oracle()
    for dt in orc.ALL_QUANT:
        assert o.lib.orc_block_bytes(dt) == orc.BYTES[dt]
        assert o.lib.orc_block_capacity(dt) == orc.CAP[dt]
        assert o.lib.orc_row_bytes(dt, 4096) == 4096 // orc.CAP[dt] * orc.BYTES[dt]


def test_moe_topk_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "moe_topk.npz"))
    scores = g["scores"].astype(np.float32)
    for k in (1, 2, 4):
        for norm in (0, 1):
            ids, wts, cnt = orc.oracle().moe_topk(scores, k, bool(norm))
            assert np.array_equal(ids, g[f"ids_k{k}_n{norm}"])
            assert np.array_equal(cnt, g[f"cnt_k{k}_n{norm}"])
            assert np.array_equal(wts, g[f"w_k{k}_n{norm}"])


@pytest.mark.skipif(not orc.have_ref_cpu(), reason="oracle/_ref not built (GPU box / fresh clone)")
@pytest.mark.parametrize("dt", orc.ALL_QUANT, ids=[orc.NAMES[d] for d in orc.ALL_QUANT])
def test_oracle_vs_reference_live(dt):
    ref = orc.RefCpu()
    assert ref.lib.ifref_sizeof_block(dt) == orc.BYTES[dt]
    assert ref.lib.ifref_byte_count(dt, 4096) == orc.row_bytes(dt, 4096)
    rng = np.random.default_rng(99 + dt)
    for sigma in (0.02, 1.0):
        src = orc.f16(rng.normal(0, sigma, (64, 512)))
        if dt == orc.Q4_B16:
            src = orc.f16(np.clip(src, -0.99, 1.4))
        a = orc.oracle().quantize(dt, src, host_q8=(dt == orc.Q8_B32T2))
        b = ref.quantize(dt, src)
        assert np.array_equal(a, b)
        assert np.array_equal(orc.oracle().dequantize(dt, a, 512), ref.dequantize(dt, b, 512))


def test_round_trip_error_bounds():
    """size-independent property: |dequant(quant(x)) - x| <= scale/2 (+fp16 rounding of params)"""
    rng = np.random.default_rng(7)
    src = orc.f16(rng.normal(0, 0.02, (16, 1024)))
    o = orc.oracle()
    steps = {orc.Q8_B32T1: 255, orc.Q6_B64T1: 62, orc.Q5_B64T1: 30, orc.Q4_B32T1A: 15,
             orc.Q4_B64T1: 14, orc.Q3H_B64T1: 10, orc.Q3_B32T1A: 7, orc.Q2_B32T1A: 3}
    for dt, div in steps.items():
        d = o.dequantize(dt, o.quantize(dt, src), 1024)
        cap = orc.CAP[dt]
        blk = src.reshape(-1, cap)
        rng_ = blk.max(1) - blk.min(1)
        err = np.abs(d.reshape(-1, cap) - blk).max(1)
        assert np.all(err <= rng_ / div * 0.5 * 1.02 + 2e-4), orc.NAMES[dt]


def test_q3h_fast_div11_exhaustive():
    """the device decoder's P/11 shortcut (ifb_formats.cuh) is exact on the 7-bit pair range"""
    assert all((P * 187) >> 11 == P // 11 for P in range(128))


def test_router_softmax_restatement_is_a_softmax():
    """tests/orc_model.py restates the reference softmax kernel's fp16 arithmetic for the MoE router: it must stay
    within half-precision distance of the exact softmax, sum to ~1 and keep the order of the logits"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from orc_model import OracleModel
    rng = np.random.default_rng(9)
    for n in (2, 4, 8, 16, 32):
        for _ in range(20):
            z = rng.normal(0, 3, n).astype(np.float16)
            p = OracleModel.router_softmax(z).astype(np.float64)
            e = np.exp(z.astype(np.float64) - z.astype(np.float64).max())
            want = e / e.sum()
            assert np.abs(p - want).max() <= 2 ** -9
            assert abs(p.sum() - 1.0) <= n * 2 ** -10
            order = np.argsort(-z.astype(np.float64), kind="stable")
            assert np.all(np.diff(p[order]) <= 1e-12)
