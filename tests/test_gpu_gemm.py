"""Dequant-fused tensor-core GEMM (csrc/ifb_gemm.cu, ifb_gemm_q: tcgen05.mma with the accumulator in tensor memory)
against the fp32 oracle the survey prescribes for the rows > 1 path (the reference dequantizes W to fp16 and
calls cuBLAS -- no reference test pins results there): Y_ref = X . dequant(W)^T in fp32, with dequant(W) the fp16 bytes of
TensorOpr::Dequantize (our ifb_dequantize_rows, bit-exact against the reference in test_gpu_ops / golden fixtures).
Tolerance: fp16 output rounding of an fp32-accumulated sum: 2^-10 of the result's range (+ the fp16 ulp of the value)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FORMATS = [orc.Q4_B32T1A, orc.Q4_B64T1, orc.Q8_B32T2, orc.Q6_B64T1, orc.Q5_B64T1, orc.Q3H_B64T1, orc.F16]
SHAPES = [(8, 256, 256), (17, 384, 512), (128, 256, 1024), (300, 512, 448), (1, 128, 64)]


def _n(dt):
    return "F16" if dt == orc.F16 else orc.NAMES[dt]


def _quantized(ib, dt, cy, cx, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    w = (torch.randn(cy, cx, device="cuda", generator=g) * 0.05).to(torch.float16)
    if dt == orc.F16:
        return w, w
    wq = torch.empty(cy * ib.row_bytes(dt, cx), dtype=torch.uint8, device="cuda")
    ib.check(ib.lib.ifb_quantize_rows(dt, wq.data_ptr(), w.data_ptr(), cy, cx, None), "quantize")
    wd = torch.empty(cy, cx, dtype=torch.float16, device="cuda")
    ib.check(ib.lib.ifb_dequantize_rows(dt, wd.data_ptr(), wq.data_ptr(), cy, cx, None), "dequantize")
    return wq, wd


@pytest.mark.parametrize("dt", FORMATS, ids=[_n(d) for d in FORMATS])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"t{t}_cy{a}_cx{b}" for t, a, b in SHAPES])
def test_gemm_q_matches_fp32_oracle(dt, shape):
    import inferflow_b200 as ib
    tokens, cy, cx = shape
    wq, wd = _quantized(ib, dt, cy, cx, seed=cy + cx)
    g = torch.Generator(device="cuda"); g.manual_seed(tokens)
    x = torch.randn(tokens, cx, device="cuda", generator=g).to(torch.float16)
    y = torch.full((tokens, cy), float("nan"), dtype=torch.float16, device="cuda")
    ib.check(ib.lib.ifb_gemm_q(dt, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, None, None), "ifb_gemm_q")
    torch.cuda.synchronize()
    want = x.float() @ wd.float().T
    err = (y.float() - want).abs().max().item()
    rng_ = want.abs().max().item()
    assert np.isfinite(err), "NaN / unwritten outputs"
    assert err <= 2 ** -10 * rng_ + 2 ** -11 * rng_, f"{_n(dt)} {shape}: |y - oracle| = {err:.3e} (range {rng_:.3e})"


def test_gemm_q_bias_and_bad_shapes():
    import inferflow_b200 as ib
    tokens, cy, cx = 40, 256, 256
    wq, wd = _quantized(ib, orc.Q4_B32T1A, cy, cx, seed=1)
    x = torch.randn(tokens, cx, device="cuda").to(torch.float16)
    b = torch.randn(cy, device="cuda").to(torch.float16)
    y = torch.empty(tokens, cy, dtype=torch.float16, device="cuda")
    ib.check(ib.lib.ifb_gemm_q(orc.Q4_B32T1A, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, b.data_ptr(), None), "gemm")
    torch.cuda.synchronize()
    want = (x.float() @ wd.float().T).to(torch.float16).float() + b.float()
    assert (y.float() - want).abs().max().item() <= 2 ** -9 * want.abs().max().item()
    assert ib.lib.ifb_gemm_q(orc.Q4_B32T1A, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, 200, cx, None, None) != 0   # cy % 128
    assert ib.lib.ifb_gemm_q(orc.Q2_B32T1B, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, None, None) != 0    # no tensor-core path
