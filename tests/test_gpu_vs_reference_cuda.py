"""Ours vs THE REFERENCE'S OWN CUDA OPERATORS on the same device buffers.

oracle/_ref/libifref_cuda.so is the reference's TensorMul / TensorOpr layer compiled natively
for sm_100a from /root/reference (oracle/Makefile + oracle/ref_cuda_shim.cu) -- a prebuilt file on
the GPU box. This pins what the CPU-side fixtures cannot: the device flavour of the Q8_B32T2
activation quantizer (alg 2), the dp4a GEMV kernels, RoPE/RMSNorm with the device math library,
and the whole attention operator sequence.

Tolerance policy: with err_ref = |reference_cuda - exact oracle|, require
|ours - exact oracle| <= max(err_ref, 2^-10 * max|oracle|); integer / byte outputs bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

if not os.path.exists(orc.ref_cuda_path()):
    pytest.skip("oracle/_ref/libifref_cuda.so not built", allow_module_level=True)


@pytest.fixture(scope="module")
def ref():
    torch.zeros(1, device="cuda")
    L = C.CDLL(orc.ref_cuda_path())
    vp, i = C.c_void_p, C.c_int
    L.ifref_cuda_gemv_ax.argtypes = [i, vp, vp, i, vp, i, i]
    L.ifref_cuda_quantize.argtypes = [i, vp, vp, i, i]
    L.ifref_cuda_dequantize.argtypes = [i, vp, vp, i, i]
    L.ifref_cuda_rmsnorm.argtypes = [vp, vp, vp, i, i]
    L.ifref_cuda_rope.argtypes = [vp, i, i, i, i, i, C.c_float]
    L.ifref_cuda_silu_mul.argtypes = [vp, vp, vp, i]
    L.ifref_cuda_attention_decode.argtypes = [vp, vp, vp, vp, i, i, i, i, C.c_float]
    return L


@pytest.fixture(scope="module")
def ib():
    import inferflow_b200
    return inferflow_b200


_KEEP = []


@pytest.fixture(autouse=True)
def _keep_alive():
    """device tensors created with dev() stay allocated until the test ends (a temporary passed as
    ptr(dev(x)) would otherwise be freed -- and its memory reused -- before the kernel runs)"""
    _KEEP.clear()
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    _KEEP.append(t)
    return t


def ptr(t):
    return C.c_void_p(t.data_ptr())


def policy(ours, ref_out, exact):
    ours, ref_out, exact = (np.asarray(a, dtype=np.float64) for a in (ours, ref_out, exact))
    e_ours, e_ref = np.abs(ours - exact).max(), np.abs(ref_out - exact).max()
    floor = 2.0 ** -10 * np.abs(exact).max()
    assert e_ours <= max(e_ref, floor) * 1.0001 + 1e-30, f"ours {e_ours:.3e} vs reference {e_ref:.3e} (floor {floor:.3e})"
    return e_ours, e_ref


@pytest.mark.parametrize("dt", orc.ALL_QUANT, ids=[orc.NAMES[d] for d in orc.ALL_QUANT])
def test_quantize_same_bytes_as_reference_cuda(ib, ref, dt):
    rng = np.random.default_rng(200 + dt)
    rows, cx = 128, 1024
    src = orc.f16(rng.normal(0, 0.05, (rows, cx)))
    if dt == orc.Q4_B16:
        src = orc.f16(np.clip(src * 10, -0.99, 1.4))   # the format's valid base range
    x = dev(src.astype(np.float16))
    nbytes = rows * orc.row_bytes(dt, cx)
    a = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    b = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    ib.check(ib.lib.ifb_quantize_rows(dt, ptr(a), ptr(x), rows, cx, None), "quantize")
    assert ref.ifref_cuda_quantize(dt, ptr(b), ptr(x), rows, cx) == 0
    torch.cuda.synchronize()
    assert torch.equal(a, b), f"{orc.NAMES[dt]}: {(a != b).sum().item()} bytes differ from the reference CUDA quantizer"


@pytest.mark.parametrize("dt", orc.AX8_TYPES, ids=[orc.NAMES[d] for d in orc.AX8_TYPES])
def test_gemv_dp4a_vs_reference_cuda(ib, ref, dt):
    cy, cx = 512, 4096
    rng = np.random.default_rng(210 + dt)
    w = orc.oracle().quantize(dt, orc.f16(rng.normal(0, 0.02, (cy, cx))))
    x = orc.f16(rng.normal(0, 1, cx))
    wt, xt = dev(w), dev(x.astype(np.float16))
    xq = torch.zeros(orc.row_bytes(orc.Q8_B32T2, cx), dtype=torch.uint8, device="cuda")
    assert ref.ifref_cuda_quantize(orc.Q8_B32T2, ptr(xq), ptr(xt), 1, cx) == 0     # the reference's own activation quantizer
    y_ref = torch.zeros(cy, dtype=torch.float16, device="cuda")
    y_our = torch.zeros(cy, dtype=torch.float16, device="cuda")
    assert ref.ifref_cuda_gemv_ax(dt, ptr(y_ref), ptr(wt), orc.Q8_B32T2, ptr(xq), cy, cx) == 0
    ib.check(ib.lib.ifb_gemv(dt, ptr(y_our), ptr(wt), orc.Q8_B32T2, ptr(xq), cy, cx, None, None), "gemv")
    torch.cuda.synchronize()
    exact = orc.oracle().gemv_ax8(dt, w, xq.cpu().numpy(), cy, cx)
    policy(y_our.cpu().numpy(), y_ref.cpu().numpy(), exact)
    # the fused path (quantize inside our kernel) sees the same activation codes
    from inferflow_b200.capi import GemvDesc
    d = GemvDesc()
    d.w_dtype = dt; d.nseg = 1; d.w[0] = wt.data_ptr(); d.rows[0] = cy; d.cx = cx
    d.x_f16 = xt.data_ptr(); d.y_f16 = y_our.data_ptr(); d.quantize_x = 1; d.norm_eps = 1e-5
    ib.check(ib.lib.ifb_gemv_fused(C.byref(d), None), "gemv_fused")
    torch.cuda.synchronize()
    policy(y_our.cpu().numpy(), y_ref.cpu().numpy(), exact)


@pytest.mark.parametrize("dt", [orc.Q3_B32T1B, orc.Q2_B32T1B, orc.Q8_B32T1, orc.Q5_B32T1, orc.Q6_B64T1, orc.F16],
                         ids=["Q3", "Q2", "Q8_B32T1", "Q5_B32T1", "Q6", "F16"])
def test_gemv_half_activation_vs_reference_cuda(ib, ref, dt):
    cy, cx = 256, 4096
    rng = np.random.default_rng(220 + dt)
    wf = orc.f16(rng.normal(0, 0.02, (cy, cx)))
    x = orc.f16(rng.normal(0, 1, cx))
    if dt == orc.F16:
        w = wf.astype(np.float16).view(np.uint8).reshape(cy, -1)
        exact = orc.oracle().gemv_f16(wf.astype(np.float16), x, cy, cx)
    else:
        w = orc.oracle().quantize(dt, wf)
        exact = orc.oracle().gemv_half(dt, w, x, cy, cx)
    wt, xt = dev(w), dev(x.astype(np.float16))
    y_ref = torch.zeros(cy, dtype=torch.float16, device="cuda")
    y_our = torch.zeros(cy, dtype=torch.float16, device="cuda")
    assert ref.ifref_cuda_gemv_ax(dt, ptr(y_ref), ptr(wt), orc.F16, ptr(xt), cy, cx) == 0
    ib.check(ib.lib.ifb_gemv(dt, ptr(y_our), ptr(wt), orc.F16, ptr(xt), cy, cx, None, None), "gemv")
    torch.cuda.synchronize()
    policy(y_our.cpu().numpy(), y_ref.cpu().numpy(), exact)


def test_rmsnorm_rope_silu_vs_reference_cuda(ib, ref):
    o = orc.oracle()
    rng = np.random.default_rng(230)
    x = orc.f16(rng.normal(0, 1, (2, 4096))); w = orc.f16(rng.normal(1, 0.1, 4096))
    xt, wt = dev(x.astype(np.float16)), dev(w.astype(np.float16))
    a = torch.zeros(2 * 4096, dtype=torch.float16, device="cuda"); b = torch.zeros_like(a)
    ib.check(ib.lib.ifb_rmsnorm(ptr(a), ptr(xt), ptr(wt), 2, 4096, 1e-5, 0.0, None), "rmsnorm")
    assert ref.ifref_cuda_rmsnorm(ptr(b), ptr(xt), ptr(wt), 2, 4096) == 0
    torch.cuda.synchronize()
    exact = np.stack([o.rmsnorm(x[i], w) for i in range(2)]).reshape(-1)
    policy(a.cpu().numpy(), b.cpu().numpy(), exact)

    for order in (0, 2):
        q = orc.f16(rng.normal(0, 1, (1, 32, 128)))
        qa, qb = dev(q.astype(np.float16)), dev(q.astype(np.float16))
        ib.check(ib.lib.ifb_rope(ptr(qa), 1, 32, 128, 0, 2047, 10000.0, order, None), "rope")
        assert ref.ifref_cuda_rope(ptr(qb), 1, 32, 128, 2047, order, 10000.0) == 0
        torch.cuda.synchronize()
        # same device math library, same formula: identical halves
        assert torch.equal(qa, qb), f"rope order {order}: {(qa != qb).sum().item()} halves differ"

    g = orc.f16(rng.normal(0, 2, 11008)); u = orc.f16(rng.normal(0, 1, 11008))
    gt, ut = dev(g.astype(np.float16)), dev(u.astype(np.float16))
    oa = torch.zeros(11008, dtype=torch.float16, device="cuda"); ob = torch.zeros_like(oa)
    ib.check(ib.lib.ifb_silu_mul(ptr(oa), ptr(gt), ptr(ut), 11008, None), "silu")
    assert ref.ifref_cuda_silu_mul(ptr(ob), ptr(gt), ptr(ut), 11008) == 0
    torch.cuda.synchronize()
    policy(oa.cpu().numpy(), ob.cpu().numpy(), o.silu_mul(g, u))


ATT = [(32, 32, 128, 300), (32, 32, 128, 2048), (64, 8, 128, 515), (12, 12, 64, 129)]


@pytest.mark.parametrize("kv_dt", [orc.F16, orc.Q8_B32T2], ids=["F16", "Q8"])
@pytest.mark.parametrize("case", ATT, ids=[f"h{h}g{g}d{d}c{c}" for h, g, d, c in ATT])
def test_attention_vs_reference_cuda(ib, ref, kv_dt, case):
    heads, kv_heads, hd, ctx = case
    max_ctx = 2048
    rng = np.random.default_rng(240 + ctx)
    k = orc.f16(rng.normal(0, 1, (ctx, kv_heads * hd))); v = orc.f16(rng.normal(0, 1, (ctx, kv_heads * hd)))
    q = orc.f16(rng.normal(0, 1, heads * hd))
    nb = ib.lib.ifb_kv_cache_bytes(kv_dt, kv_heads, hd, max_ctx)
    kc = torch.zeros(nb, dtype=torch.uint8, device="cuda"); vc = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    kt, vt, qt = dev(k.astype(np.float16)), dev(v.astype(np.float16)), dev(q.astype(np.float16))
    ib.check(ib.lib.ifb_kv_append(kv_dt, ptr(kc), ptr(vc), ptr(kt), ptr(vt), kv_heads, hd, max_ctx, 0, ctx, None), "append")
    # what the reference's GetKRows/GetVRows would hand to its attention: our cache read back as fp16 rows
    kr = torch.zeros(ctx * kv_heads * hd, dtype=torch.float16, device="cuda"); vr = torch.zeros_like(kr)
    ib.check(ib.lib.ifb_kv_get_rows(kv_dt, ptr(kr), ptr(kc), kv_heads, hd, max_ctx, 0, ctx, None), "get")
    ib.check(ib.lib.ifb_kv_get_rows(kv_dt, ptr(vr), ptr(vc), kv_heads, hd, max_ctx, 0, ctx, None), "get")
    if kv_dt == orc.Q8_B32T2:
        # and those rows are what the reference's own Quantize -> Dequantize produces
        qb = torch.zeros(ctx * orc.row_bytes(orc.Q8_B32T2, kv_heads * hd), dtype=torch.uint8, device="cuda")
        back = torch.zeros_like(kr)
        assert ref.ifref_cuda_quantize(orc.Q8_B32T2, ptr(qb), ptr(kt), ctx, kv_heads * hd) == 0
        assert ref.ifref_cuda_dequantize(orc.Q8_B32T2, ptr(back), ptr(qb), ctx, kv_heads * hd) == 0
        torch.cuda.synchronize()
        assert torch.equal(back, kr), "KV rows differ from the reference's Quantize->Dequantize"
    out_ref = torch.zeros(heads * hd, dtype=torch.float16, device="cuda")
    out_our = torch.zeros(heads * hd, dtype=torch.float16, device="cuda")
    assert ref.ifref_cuda_attention_decode(ptr(out_ref), ptr(qt), ptr(kr), ptr(vr), ctx, heads, kv_heads, hd, 1.0) == 0
    ws = torch.zeros(ib.lib.ifb_attn_workspace_bytes(heads, hd), dtype=torch.uint8, device="cuda")
    ib.check(ib.lib.ifb_attn_decode(kv_dt, ptr(out_our), ptr(qt), ptr(kc), ptr(vc), heads, kv_heads, hd, max_ctx,
                                    ctx, 1.0, ptr(ws), None), "attn")
    torch.cuda.synchronize()
    kd = kr.cpu().numpy().astype(np.float32).reshape(ctx, -1); vd = vr.cpu().numpy().astype(np.float32).reshape(ctx, -1)
    exact = orc.oracle().attn_decode(q, kd, vd, heads, kv_heads, hd, 1.0, exact=True)
    policy(out_our.cpu().numpy(), out_ref.cpu().numpy(), exact)
