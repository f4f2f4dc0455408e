"""The drop-in boundary, RUN: the reference's own CUDA engine (InferenceEngine -> GpuInferenceWorker -> TensorMul /
TensorOpr, compiled unmodified from /root/reference by oracle/Makefile) generates tokens twice from the same config --
once stock, once with the bodies of TensorMul::Gemv_AX and TensorOpr::Quantize replaced at LINK time by
integration/reference_shim.cu, which forwards to ifb_gemv / ifb_quantize_rows of libinferflow_b200.so (target
`boundary` of oracle/Makefile: the stock tensor_mul.cu / tensor_opr.cu are compiled with the two member names renamed,
no reference source is edited). Same tokens, logits within the model tolerance."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.gpu
REF = os.path.join(HERE, "..", "oracle", "_ref")
STOCK, SHIM = os.path.join(REF, "libifref_engine_stock.so"), os.path.join(REF, "libifref_engine_ifb.so")


@pytest.mark.skipif(not (os.path.exists(STOCK) and os.path.exists(SHIM)),
                    reason="oracle/_ref/libifref_engine_*.so not built (make -C oracle boundary needs /root/reference)")
@pytest.mark.parametrize("wtype", ["Q8", "Q4", "Q3H"])
def test_reference_cuda_engine_with_our_gemv_and_quantize_linked_in(wtype):
    pytest.importorskip("torch")
    import ref_engine as R
    hp = dict(dim=512, hidden=1408, layers=4, heads=8, kv_heads=8, vocab=1024, seq=128)
    w = R.synthetic_llama(hp, 77)
    prompt = [1, 17, 93, 400, 7, 256]
    with tempfile.TemporaryDirectory() as d:
        ini = R.write_tree(d, hp, w, threads=2)
        text = open(ini).read()
        # every layer on the GPU, quantized weights whatever their size (tensor_quant_threshold = 0), dp4a GEMV
        text = text.replace(f"decoder_cpu_layer_count = {hp['layers']}", "decoder_cpu_layer_count = 0")
        text = text.replace("device_weight_data_type = F16", f"device_weight_data_type = {wtype}\ntensor_quant_threshold = 0")
        open(ini, "w").write(text)
        res = {}
        for name, lib in (("stock", STOCK), ("shim", SHIM)):
            out = os.path.join(d, name + ".npz")
            p = subprocess.run([sys.executable, os.path.join(HERE, "boundary_runner.py"), lib, ini, out, str(hp["vocab"]), "16"] +
                               [str(t) for t in prompt], capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, f"{name} engine failed:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
            res[name] = np.load(out)
    a, b = res["stock"], res["shim"]
    assert len(a["tokens"]) == 16
    rng_ = np.abs(a["logits"]).max()
    assert rng_ > 0, "the engine returned no logits (return_output_tensors)"
    err = np.abs(a["logits"] - b["logits"]).max()
    # Two implementations of the same GEMV (ours accumulates in fp32 where the reference's kernels accumulate partly in
    # half, gemv.h) inside an engine that re-quantizes every activation to Q8_B32T2: a last-bit difference can move an
    # int8 code, so the END-TO-END logits agree less tightly than one operator does (2^-7, tests/test_gpu_vs_reference_
    # cuda.py). Measured over 4 layers: Q8 weights 2^-9, Q4 2^-6.3 of the logit range. Bound: 2^-5; tokens must be the
    # same wherever the stock engine's top-2 margin exceeds 4x the difference.
    assert err <= 2 ** -5 * rng_, f"logits differ by {err:.3e} (range {rng_:.2f})"
    for k in range(16):
        srt = np.sort(a["logits"][k])
        if srt[-1] - srt[-2] > 4 * np.abs(a["logits"][k] - b["logits"][k]).max():
            assert a["tokens"][k] == b["tokens"][k], f"step {k}"
        else:
            break       # a near-tie may legitimately fork the two streams; later steps see different prefixes
