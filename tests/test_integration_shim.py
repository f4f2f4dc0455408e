"""INTEGRATION.md is not prose only: integration/reference_shim.cu -- the reference-side binding of the operator
seam (TensorMul::Gemv_AX, TensorOpr::Quantize/Dequantize/RmsNorm, decode attention) over our C ABI -- must compile
against the reference's OWN headers, and every C-ABI function it calls must be one the library exports. Needs
/root/reference (authoring container); skipped elsewhere."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")) or not os.path.exists(NVCC),
                    reason="needs the reference sources and nvcc")
def test_reference_side_shim_compiles_against_reference_headers():
    src = os.path.join(ROOT, "integration", "reference_shim.cu")
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "shim.o")
        cmd = [NVCC, "-O2", "-std=c++14", "-DUSE_CUDA", "-DHAVE_CONFIG_H", "-DNDEBUG", "-w",
               "-gencode", "arch=compute_100a,code=sm_100a", f"-I{REF}/src", f"-I{REF}/3rd_party",
               f"-I{ROOT}/include", "-c", src, "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        nm = shutil.which("nm")
        if nm:
            syms = subprocess.run([nm, "-u", out], capture_output=True, text=True).stdout
            used = set(re.findall(r"\bifb_[a-z0-9_]+\b", syms))
            assert {"ifb_gemv", "ifb_quantize_rows", "ifb_rmsnorm", "ifb_attn_decode"} <= used


def test_shim_calls_only_exported_functions():
    import inferflow_b200 as ib
    text = open(os.path.join(ROOT, "integration", "reference_shim.cu")).read()
    called = set(re.findall(r"\b(ifb_[a-z0-9_]+)\s*\(", text))
    assert called, "the shim should call the C ABI"
    for name in called:
        assert name in ib.capi.DECLARED and hasattr(ib.lib, name), f"{name} is not exported by the library"
