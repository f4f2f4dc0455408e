"""The work plan of the persistent decode-step kernel (csrc/ifb_step.cu: step_build), computed on the host through
ifb_step_plan -- no GPU: every row of every matrix is covered exactly once, units and chunks respect the 16-byte
granularity of TMA bulk copies and fit their shared-memory slots, the shared-memory budget of an SM is kept, and models
the warp-private schedule cannot take fall back to the shared ring."""
import ctypes as C

import pytest

import orc

KINDS = ["QKV", "ATT", "WO", "GLU", "W2", "HEAD"]
SMEM_MAX = 232448 - 1536
PRIV_CHUNK = 5120


def _plan(ib, spec, priv):
    out = (C.c_int32 * 128)()
    c = spec.to_c()
    rc = ib.lib.ifb_step_plan(C.byref(c), 1 if priv else 0, out, 128)
    if rc != 0:
        return None
    v = list(out[:84])
    ops = {}
    for i, k in enumerate(KINDS):
        r = v[i * 13:(i + 1) * 13]
        ops[k] = dict(zip(["units", "tasks", "R", "cx", "nblk", "row_bytes", "unit_bytes", "vec4", "rt", "nc", "kc_blk", "kc_bytes", "ntasks"], r))
    tail = dict(zip(["priv", "nslots", "slot_bytes", "smem_bytes", "att_pos", "pslot"], v[78:84]))
    return ops, tail


def _specs(ib):
    base7, base70 = ib.LLAMA2_7B, ib.LLAMA2_70B
    out = []
    for tp in (1, 2, 4, 8):
        out.append(("7B-tp%d" % tp, ib.ModelSpec(max_ctx=2048, tp_size=tp, **base7)))
    out.append(("70B-tp8", ib.ModelSpec(max_ctx=2048, tp_size=8, **base70)))
    out.append(("70B-tp1", ib.ModelSpec(max_ctx=2048, **base70)))
    out.append(("7B-q3h-repacked", ib.ModelSpec(max_ctx=2048, weight_dtype=orc.Q3H_B64T1, weight_repack=1, **base7)))
    for dt in (orc.Q4_B32T1A, orc.Q4_B64T1, orc.Q8_B32T2, orc.Q6_B64T1, orc.Q5_B64T1, orc.Q3H_B64T1, orc.F16):
        ffn = 1376 if dt in (orc.Q4_B32T1A, orc.Q8_B32T2, orc.F16) else 1408      # 64-element blocks need d_ffn % 64 == 0
        out.append(("small-%d" % dt, ib.ModelSpec(layers=2, d_model=512, d_ffn=ffn, heads=8, kv_heads=2, vocab=512, max_ctx=128,
                                               weight_dtype=dt, kv_dtype=orc.Q8_B32T2)))
        out.append(("long-rows-%d" % dt, ib.ModelSpec(layers=2, d_model=2048, d_ffn=5504, heads=16, kv_heads=16, vocab=1024, max_ctx=128,
                                                   weight_dtype=dt, kv_dtype=orc.Q8_B32T2, lm_head_dtype=dt)))
    return out


def _rows(spec):
    tp = max(1, spec.tp_size)
    hd = spec.d_model // spec.heads
    q, kv = spec.heads // tp * hd, spec.kv_heads // tp * hd
    return {"QKV": q + 2 * kv, "WO": spec.d_model, "GLU": spec.d_ffn // tp, "W2": spec.d_model, "HEAD": spec.vocab // tp}


@pytest.mark.parametrize("priv", [False, True], ids=["ring", "private"])
def test_step_plan_invariants(priv):
    import inferflow_b200 as ib
    seen_priv = seen_fallback = 0
    for name, spec in _specs(ib):
        p = _plan(ib, spec, priv)
        assert p is not None, f"{name}: the step kernel should cover this model ({ib.lib.ifb_last_error()})"
        ops, tail = p
        rows = _rows(spec)
        assert tail["smem_bytes"] <= SMEM_MAX, name
        for k, total in rows.items():
            o = ops[k]
            # the shared-ring geometry is always there (the fallback of the private schedule)
            assert o["units"] * o["R"] == total and o["R"] % 2 == 0, (name, k, o)
            assert o["unit_bytes"] == o["R"] * o["row_bytes"] and o["unit_bytes"] % 16 == 0, (name, k, o)
        if not tail["priv"]:
            assert tail["nslots"] >= 3
            assert tail["slot_bytes"] >= max(ops["GLU"]["unit_bytes"] * 2, ops["QKV"]["unit_bytes"], ops["W2"]["unit_bytes"], ops["HEAD"]["unit_bytes"]), name
            seen_fallback += int(priv)
            continue
        seen_priv += 1
        assert tail["nslots"] == 30 and tail["pslot"] >= PRIV_CHUNK + 16 and tail["att_pos"] == 32, name
        for k, total in rows.items():
            o = ops[k]
            assert o["rt"] in (2, 4) and o["ntasks"] * o["rt"] == total, (name, k, o)
            assert (o["rt"] * o["row_bytes"]) % 16 == 0, (name, k, o)
            if o["nc"] == 1:
                assert o["rt"] * o["row_bytes"] <= PRIV_CHUNK, (name, k, o)
            else:
                assert o["rt"] == 2 and o["row_bytes"] % 16 == 0, (name, k, o)
                assert o["kc_bytes"] % 16 == 0 and 2 * o["kc_bytes"] <= PRIV_CHUNK, (name, k, o)
                f16 = o["row_bytes"] == o["cx"] * 2
                n_el = o["cx"] if f16 else o["nblk"]             # a chunk is counted in halves (fp16 rows) or blocks
                assert (o["nc"] - 1) * o["kc_blk"] < n_el <= o["nc"] * o["kc_blk"], (name, k, o)
                assert o["kc_bytes"] * n_el == o["kc_blk"] * o["row_bytes"], (name, k, o)    # bytes per chunk = whole blocks
    if priv:
        assert seen_priv >= 10 and seen_fallback >= 1       # e.g. 70B on one GPU: the staging areas leave no room for 30 slots


def test_step_plan_refuses_what_the_kernel_does_not_cover():
    import inferflow_b200 as ib
    out = (C.c_int32 * 128)()
    for spec in (ib.ModelSpec(layers=2, d_model=256, d_ffn=512, heads=8, kv_heads=8, vocab=512, max_ctx=64),            # head_dim 32
                 ib.ModelSpec(layers=2, d_model=512, d_ffn=1024, heads=4, kv_heads=4, vocab=512, max_ctx=64, weight_dtype=orc.Q2_B32T1A)):
        c = spec.to_c()
        assert ib.lib.ifb_step_plan(C.byref(c), 0, out, 128) != 0
        assert b"step kernel" in ib.lib.ifb_last_error()
