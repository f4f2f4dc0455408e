"""inferflow_b200 -- B200-native decode hot path of inferflow/inferflow.

The product is the C-ABI shared library `libinferflow_b200.so` (include/inferflow_b200.h),
built from inferflow_b200/csrc/*.cu for sm_100a. This package is only the ctypes door to
it plus small helpers for tests and bench.py; there is NO Python / CPU fallback: if the
library is missing or there is no CUDA device, calls raise.
"""
from .capi import lib, check, IfbError, DT, dtype_from_name, row_bytes  # noqa: F401
from .model import Model, ModelSpec, Engine, resolve_config, probe_llama2c, LLAMA2_7B, LLAMA2_70B, STORIES_110M  # noqa: F401
