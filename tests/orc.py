"""ctypes doors to the checkers (TEST INFRASTRUCTURE ONLY).

`oracle`  -> oracle/liboracle.so          (our plain-C restatement, oracle/oracle.c)
`ref_cpu` -> oracle/_ref/libifref_cpu.so  (the reference itself, CPU build; optional)
`ref_cuda`-> oracle/_ref/libifref_cuda.so (the reference's CUDA operator layer, sm_100a; GPU box)

The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# reference ElementType ids (src/tensor/tensor_common.h:15-42)
F32, F16 = 0, 1
Q8_B32T1, Q8_B32T2, Q6_B64T1, Q5_B64T1, Q5_B32T1, Q4_B16 = 7, 8, 9, 10, 11, 12
Q4_B32T1A, Q4_B32T1B, Q4_B64T1, Q3H_B64T1 = 13, 14, 17, 18
Q3_B32T1A, Q3_B32T1B, Q2_B32T1A, Q2_B32T1B = 19, 20, 21, 22

NAMES = {
    Q8_B32T1: "Q8_B32T1", Q8_B32T2: "Q8_B32T2", Q6_B64T1: "Q6_B64T1", Q5_B64T1: "Q5_B64T1",
    Q5_B32T1: "Q5_B32T1", Q4_B16: "Q4_B16", Q4_B32T1A: "Q4_B32T1A", Q4_B32T1B: "Q4_B32T1B",
    Q4_B64T1: "Q4_B64T1", Q3H_B64T1: "Q3H_B64T1", Q3_B32T1A: "Q3_B32T1A", Q3_B32T1B: "Q3_B32T1B",
    Q2_B32T1A: "Q2_B32T1A", Q2_B32T1B: "Q2_B32T1B",
}
ALL_QUANT = sorted(NAMES)
# weight types that take the dp4a path (inference_worker.cc:2707-2730)
AX8_TYPES = [Q8_B32T2, Q6_B64T1, Q5_B64T1, Q4_B32T1A, Q4_B32T1B, Q4_B64T1, Q3H_B64T1]

CAP = {Q8_B32T1: 32, Q8_B32T2: 32, Q6_B64T1: 64, Q5_B64T1: 64, Q5_B32T1: 32, Q4_B16: 16,
       Q4_B32T1A: 32, Q4_B32T1B: 32, Q4_B64T1: 64, Q3H_B64T1: 64, Q3_B32T1A: 32, Q3_B32T1B: 32,
       Q2_B32T1A: 32, Q2_B32T1B: 32}
BYTES = {Q8_B32T1: 36, Q8_B32T2: 34, Q6_B64T1: 52, Q5_B64T1: 44, Q5_B32T1: 24, Q4_B16: 10,
         Q4_B32T1A: 20, Q4_B32T1B: 20, Q4_B64T1: 36, Q3H_B64T1: 32, Q3_B32T1A: 16,
         Q3_B32T1B: 16, Q2_B32T1A: 12, Q2_B32T1B: 12}


def row_bytes(dt, cx):
    return cx // CAP[dt] * BYTES[dt]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_row_bytes.restype = C.c_int64
        L.orc_row_bytes.argtypes = [C.c_int, C.c_int64]
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float]
        L.orc_rope.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_attn_decode.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_float, C.c_int]
        L.orc_moe_topk.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_float]
        for f in (L.orc_quantize_row, L.orc_dequantize_row):
            f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_quantize_row_q8t2_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_gemv_ax8.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_gemv_half.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_gemv_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_silu_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_block_codes.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    # src: float32 [rows, cx] -> uint8 [rows, row_bytes]
    def quantize(self, dt, src, host_q8=False):
        src = np.ascontiguousarray(src, dtype=np.float32)
        rows, cx = src.reshape(-1, src.shape[-1]).shape
        out = np.zeros((rows, row_bytes(dt, cx)), dtype=np.uint8)
        if host_q8:
            assert dt == Q8_B32T2
            rc = self.lib.orc_quantize_row_q8t2_host(out.ctypes.data, src.ctypes.data, rows * cx)
        else:
            rc = self.lib.orc_quantize_row(dt, out.ctypes.data, src.ctypes.data, rows * cx)
        assert rc == 0
        return out

    def dequantize(self, dt, blocks, cx):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        rows = blocks.size // row_bytes(dt, cx)
        out = np.zeros((rows, cx), dtype=np.float32)
        assert self.lib.orc_dequantize_row(dt, out.ctypes.data, blocks.ctypes.data, rows * cx) == 0
        return out

    def block_codes(self, dt, blocks, cx):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
        n = blocks.size // BYTES[dt]
        codes = np.zeros((n, CAP[dt]), dtype=np.int32)
        sc = np.zeros(n, dtype=np.float32)
        ba = np.zeros(n, dtype=np.float32)
        tmp = np.zeros(64, dtype=np.int32)
        s1 = C.c_float()
        b1 = C.c_float()
        for i in range(n):
            blk = blocks[i * BYTES[dt]:(i + 1) * BYTES[dt]]
            self.lib.orc_block_codes(dt, tmp.ctypes.data, C.byref(s1), C.byref(b1), blk.ctypes.data)
            codes[i] = tmp[:CAP[dt]]
            sc[i], ba[i] = s1.value, b1.value
        return codes, sc, ba

    def gemv_ax8(self, dt, w, xq, cy, cx):
        y = np.zeros(cy, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.uint8)
        xq = np.ascontiguousarray(xq, dtype=np.uint8)
        assert self.lib.orc_gemv_ax8(dt, y.ctypes.data, w.ctypes.data, xq.ctypes.data, cy, cx) == 0
        return y

    def gemv_half(self, dt, w, x, cy, cx):
        y = np.zeros(cy, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert self.lib.orc_gemv_half(dt, y.ctypes.data, w.ctypes.data, x.ctypes.data, cy, cx) == 0
        return y

    def gemv_f16(self, w_f16, x, cy, cx):
        y = np.zeros(cy, dtype=np.float64)
        w = np.ascontiguousarray(w_f16, dtype=np.float16)
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert self.lib.orc_gemv_f16(y.ctypes.data, w.ctypes.data, x.ctypes.data, cy, cx) == 0
        return y

    def rmsnorm(self, x, w, eps=1e-5, multi_base=0.0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.zeros_like(x)
        wp = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        self.lib.orc_rmsnorm(y.ctypes.data, x.ctypes.data, None if wp is None else wp.ctypes.data,
                             x.size, eps, multi_base)
        return y

    def rope(self, x, heads, head_dim, pos, theta=10000.0, order=2, rope_dims=0):
        x = np.array(x, dtype=np.float32, copy=True).reshape(heads, head_dim)
        self.lib.orc_rope(x.ctypes.data, heads, head_dim, rope_dims, pos, theta, order)
        return x

    def silu_mul(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        out = np.zeros_like(a)
        self.lib.orc_silu_mul(out.ctypes.data, a.ctypes.data, b.ctypes.data, a.size)
        return out

    def attn_decode(self, q, k, v, heads, kv_heads, head_dim, kq_scale=1.0, exact=True):
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = np.ascontiguousarray(k, dtype=np.float32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        ctx = k.shape[0]
        out = np.zeros(heads * head_dim, dtype=np.float32)
        self.lib.orc_attn_decode(out.ctypes.data, q.ctypes.data, k.ctypes.data, v.ctypes.data, ctx,
                                 heads, kv_heads, head_dim, kq_scale, 1 if exact else 0)
        return out

    def moe_topk(self, scores, top_k, norm=True, min_score=1e-5):
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        rows, experts = scores.shape
        ids = np.zeros((rows, top_k), dtype=np.int32)
        wts = np.zeros((rows, top_k), dtype=np.float32)
        cnt = np.zeros(rows, dtype=np.int32)
        self.lib.orc_moe_topk(ids.ctypes.data, wts.ctypes.data, cnt.ctypes.data, scores.ctypes.data,
                              rows, experts, top_k, 1 if norm else 0, min_score)
        return ids, wts, cnt


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def ref_cpu_path():
    return os.path.join(ORACLE_DIR, "_ref", "libifref_cpu.so")


def ref_cuda_path():
    return os.path.join(ORACLE_DIR, "_ref", "libifref_cuda.so")


class RefCpu:
    """The reference itself (CPU build)."""

    def __init__(self):
        self.lib = C.CDLL(ref_cpu_path())
        L = self.lib
        L.ifref_byte_count.restype = C.c_longlong
        L.ifref_byte_count.argtypes = [C.c_int, C.c_longlong]
        L.ifref_quantize_row.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ifref_dequantize_row.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ifref_get_int4_row.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ifref_moe_topk.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4
        L.ifref_engine_create.restype = C.c_void_p
        L.ifref_engine_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.ifref_engine_destroy.argtypes = [C.c_void_p]
        L.ifref_engine_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p]

    def quantize(self, dt, src):
        src = np.ascontiguousarray(src, dtype=np.float32)
        rows, cx = src.reshape(-1, src.shape[-1]).shape
        out = np.zeros((rows, row_bytes(dt, cx)), dtype=np.uint8)
        for r in range(rows):
            assert self.lib.ifref_quantize_row(dt, out[r].ctypes.data, src.reshape(rows, cx)[r].ctypes.data, cx) == 0
        return out

    def dequantize(self, dt, blocks, cx):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        rows = blocks.size // row_bytes(dt, cx)
        out = np.zeros((rows, cx), dtype=np.float32)
        b = blocks.reshape(rows, -1)
        for r in range(rows):
            assert self.lib.ifref_dequantize_row(dt, out[r].ctypes.data, b[r].ctypes.data, cx) == 0
        return out

    def get_int4(self, dt, blocks, cx):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        rows = blocks.size // row_bytes(dt, cx)
        out = np.zeros((rows, cx // 4), dtype=np.int32)
        b = blocks.reshape(rows, -1)
        for r in range(rows):
            assert self.lib.ifref_get_int4_row(dt, out[r].ctypes.data, b[r].ctypes.data, cx) == 0
        return out

    def moe_topk(self, scores_f16, top_k, norm=True):
        s = np.ascontiguousarray(scores_f16, dtype=np.float16)
        rows, experts = s.shape
        ids = np.zeros((rows, top_k), dtype=np.int32)
        wts = np.zeros((rows, top_k), dtype=np.float32)
        cnt = np.zeros(rows, dtype=np.int32)
        assert self.lib.ifref_moe_topk(ids.ctypes.data, wts.ctypes.data, cnt.ctypes.data, s.ctypes.data,
                                       rows, experts, top_k, 1 if norm else 0) == 0
        return ids, wts, cnt


def have_ref_cpu():
    return os.path.exists(ref_cpu_path())


def f16(a):
    """round a float array through fp16 (the device stores activations as half)"""
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)
