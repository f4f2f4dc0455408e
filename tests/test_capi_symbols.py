"""CPU-side checks of the boundary: the library loads, exports every symbol the header
declares, and its pure-host entry points agree with the oracle (no compute calls)."""
import os
import re

import orc
from inferflow_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "inferflow_b200.h")).read()
    declared = set(re.findall(r"\b(ifb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(capi.lib, s)]
    assert not missing, f"declared but not exported: {missing}"
    assert declared == set(capi.DECLARED)


def test_block_geometry_matches_oracle():
    o = orc.oracle()
    for dt in orc.ALL_QUANT:
        assert capi.lib.ifb_block_bytes(dt) == o.lib.orc_block_bytes(dt)
        assert capi.lib.ifb_block_capacity(dt) == o.lib.orc_block_capacity(dt)
        for cx in (64, 4096, 11008):
            if cx % orc.CAP[dt] == 0:
                assert capi.row_bytes(dt, cx) == o.lib.orc_row_bytes(dt, cx)


def test_type_names():
    # tensor_common.cc:170-204
    m = {"q4": 13, "Q4": 13, "q3h": 18, "q8": 8, "q6": 9, "q5": 10, "q3": 20, "q2": 22,
         "q4_b64t1": 17, "f16": 1, "q4_b32t1b": 14}
    for k, v in m.items():
        assert capi.dtype_from_name(k) == v
    assert capi.dtype_from_name("nope") == -1


def test_errors_are_reported_not_swallowed():
    # argument validation happens before any CUDA call
    rc = capi.lib.ifb_gemv(13, None, None, 1, None, 8, 64, None, None)
    assert rc == -1 and b"null" in capi.lib.ifb_last_error()
    rc = capi.lib.ifb_quantize_rows(13, 1, 1, 1, 33, None)
    assert rc == -1 and b"multiple of 32" in capi.lib.ifb_last_error()
