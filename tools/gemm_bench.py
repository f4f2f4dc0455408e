"""Prefill / batched GEMM on the Llama-2-7B matrices: the dequant-fused tcgen05 kernel (ifb_gemm_q) against the
reference's recipe for rows > 1 on the same box -- dequantize the whole matrix to fp16, cuBLAS fp16 GEMM (torch.matmul),
transpose (inference_worker.cc:2364-2432). TFLOP/s = 2 * tokens * cy * cx / time; CUDA events, after warm-up."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import inferflow_b200 as ib
import orc

def ev(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
for dt in (orc.Q4_B32T1A, orc.Q3H_B64T1):
    for tokens in (8, 128, 1024, 4096):
        for cy, cx in ((4096, 4096), (11008, 4096), (4096, 11008)):
            w = (torch.randn(cy, cx, device="cuda") * 0.05).to(torch.float16)
            wq = torch.empty(cy * ib.row_bytes(dt, cx), dtype=torch.uint8, device="cuda")
            ib.check(ib.lib.ifb_quantize_rows(dt, wq.data_ptr(), w.data_ptr(), cy, cx, None), "q")
            wd = torch.empty(cy, cx, dtype=torch.float16, device="cuda")
            x = torch.randn(tokens, cx, device="cuda").to(torch.float16)
            y = torch.empty(tokens, cy, dtype=torch.float16, device="cuda")
            t_ours = ev(lambda: ib.lib.ifb_gemm_q(dt, y.data_ptr(), wq.data_ptr(), x.data_ptr(), tokens, cy, cx, None, None))
            def ref():
                ib.lib.ifb_dequantize_rows(dt, wd.data_ptr(), wq.data_ptr(), cy, cx, None)
                yt = torch.matmul(wd, x.T)          # [cy][tokens], as cublasGemmEx leaves it
                return yt.T.contiguous()            # Transpose
            t_ref = ev(ref)
            ib.lib.ifb_dequantize_rows(dt, wd.data_ptr(), wq.data_ptr(), cy, cx, None)
            err = (y.float() - x.float() @ wd.float().T).abs().max().item()
            fl = 2.0 * tokens * cy * cx
            rec = {"w": orc.NAMES[dt], "tokens": tokens, "cy": cy, "cx": cx, "ours_ms": t_ours, "ours_TFLOPs": fl / t_ours / 1e9,
                   "dequant_cublas_transpose_ms": t_ref, "ref_TFLOPs": fl / t_ref / 1e9, "speedup": t_ref / t_ours, "max_err": err}
            out.append(rec); print(json.dumps(rec), flush=True)
            del w, wq, wd
