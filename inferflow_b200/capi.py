"""ctypes binding of include/inferflow_b200.h (the same stub a maintainer of the reference
would write for any FFI)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libinferflow_b200.so")


class IfbError(RuntimeError):
    pass


class DT:
    F32, F16, BF16 = 0, 1, 2
    Q8_B32T1, Q8_B32T2, Q6_B64T1, Q5_B64T1, Q5_B32T1, Q4_B16 = 7, 8, 9, 10, 11, 12
    Q4_B32T1A, Q4_B32T1B, Q4_B64T1, Q3H_B64T1 = 13, 14, 17, 18
    Q3_B32T1A, Q3_B32T1B, Q2_B32T1A, Q2_B32T1B = 19, 20, 21, 22


class GemvDesc(C.Structure):
    _fields_ = [
        ("w_dtype", C.c_int32), ("nseg", C.c_int32),
        ("w", C.c_void_p * 3), ("rows", C.c_int32 * 3),
        ("w_gate", C.c_void_p), ("cx", C.c_int64),
        ("x_f16", C.c_void_p), ("norm_w_f16", C.c_void_p),
        ("norm_eps", C.c_float), ("norm_base", C.c_float),
        ("residual_f16", C.c_void_p), ("bias_f16", C.c_void_p), ("y_f16", C.c_void_p),
        ("quantize_x", C.c_int32), ("pdl", C.c_int32),
    ]


class ModelSpecC(C.Structure):
    _fields_ = [
        ("layers", C.c_int32), ("d_model", C.c_int32), ("d_ffn", C.c_int32), ("heads", C.c_int32),
        ("kv_heads", C.c_int32), ("vocab", C.c_int32), ("max_ctx", C.c_int32),
        ("weight_dtype", C.c_int32), ("kv_dtype", C.c_int32), ("lm_head_dtype", C.c_int32),
        ("rope_order", C.c_int32), ("rope_theta", C.c_float), ("norm_eps", C.c_float),
        ("kq_scale", C.c_float), ("tp_rank", C.c_int32), ("tp_size", C.c_int32),
        ("experts", C.c_int32), ("experts_top_k", C.c_int32), ("moe_norm_top_k_prob", C.c_int32),
        ("weight_repack", C.c_int32), ("kv_slots", C.c_int32), ("layer_begin", C.c_int32), ("layer_end", C.c_int32),
    ]


def _load():
    if not os.path.exists(_SO):
        raise IfbError(
            f"{_SO} is missing: build it with `python inferflow_b200/build.py` "
            "(or __graft_entry__.build()). There is no fallback path.")
    L = C.CDLL(_SO)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.ifb_version.restype = i32
    L.ifb_last_error.restype = C.c_char_p
    L.ifb_block_capacity.argtypes = [i32]
    L.ifb_block_bytes.argtypes = [i32]
    L.ifb_row_bytes.restype = i64
    L.ifb_row_bytes.argtypes = [i32, i64]
    L.ifb_dtype_from_name.argtypes = [C.c_char_p]
    L.ifb_memcpy_d2h.argtypes = [vp, vp, i64]
    L.ifb_memcpy_d2d.argtypes = [vp, vp, i64]
    L.ifb_memcpy_h2d.argtypes = [vp, vp, i64]
    L.ifb_memcpy_d2h_async.argtypes = [vp, vp, i64, vp]
    L.ifb_memcpy_h2d_async.argtypes = [vp, vp, i64, vp]
    L.ifb_quantize_rows.argtypes = [i32, vp, vp, i64, i64, vp]
    L.ifb_dequantize_rows.argtypes = [i32, vp, vp, i64, i64, vp]
    L.ifb_gemv.argtypes = [i32, vp, vp, i32, vp, i64, i64, vp, vp]
    L.ifb_gemv_fused.argtypes = [C.POINTER(GemvDesc), vp]
    L.ifb_gemm_q.argtypes = [i32, vp, vp, vp, i64, i64, i64, vp, vp]
    L.ifb_rmsnorm.argtypes = [vp, vp, vp, i64, i64, f32, f32, vp]
    L.ifb_rope.argtypes = [vp, i32, i32, i32, i32, i32, f32, i32, vp]
    L.ifb_silu_mul.argtypes = [vp, vp, vp, i64, vp]
    L.ifb_add.argtypes = [vp, vp, vp, i64, vp]
    L.ifb_argmax.argtypes = [vp, vp, i64, vp]
    L.ifb_kv_cache_bytes.restype = i64
    L.ifb_kv_cache_bytes.argtypes = [i32, i32, i32, i32]
    L.ifb_kv_append.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.ifb_kv_get_rows.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, vp]
    L.ifb_attn_workspace_bytes.restype = i64
    L.ifb_attn_workspace_bytes.argtypes = [i32, i32]
    L.ifb_attn_decode.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp]
    L.ifb_moe_route.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.ifb_moe_combine.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp]
    L.ifb_attn_decode_append.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, i32, f32, vp, vp]
    L.ifb_moe_topk.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.ifb_model_create.argtypes = [C.POINTER(vp), C.POINTER(ModelSpecC)]
    L.ifb_model_destroy.argtypes = [vp]
    L.ifb_model_destroy.restype = None
    L.ifb_model_weight_bytes.restype = i64
    L.ifb_model_weight_bytes.argtypes = [vp]
    L.ifb_model_kv_bytes_per_token.restype = i64
    L.ifb_model_kv_bytes_per_token.argtypes = [vp, i32]
    L.ifb_model_init_synthetic.argtypes = [vp, C.c_uint64, vp]
    L.ifb_model_load_tensor.argtypes = [vp, C.c_char_p, vp, i32, i64, i64, vp]
    L.ifb_model_tensor_ptr.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i64)]
    L.ifb_model_kv_reset.argtypes = [vp, i32, i32, C.c_uint64, vp]
    L.ifb_model_decode.argtypes = [vp, i32, vp, i32, vp]
    L.ifb_model_decode_next.argtypes = [vp, vp]
    L.ifb_model_prefill.argtypes = [vp, vp, i32, i32, vp]
    L.ifb_model_prefill_supported.argtypes = [vp]
    L.ifb_attn_prefill.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.ifb_model_issue_parts.argtypes = [vp, i32, vp]
    L.ifb_model_logits.restype = vp
    L.ifb_model_logits.argtypes = [vp]
    L.ifb_model_next_token.restype = vp
    L.ifb_model_next_token.argtypes = [vp]
    L.ifb_model_kernel_launches_per_step.argtypes = [vp]
    L.ifb_model_step_status.argtypes = [vp, vp]
    L.ifb_model_storage_dtype.argtypes = [vp, i32]
    L.ifb_step_plan.argtypes = [vp, i32, vp, i32]
    L.ifb_model_batch_supported.argtypes = [vp]
    L.ifb_model_decode_batch.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ifb_model_batch_logits.argtypes = [vp]
    L.ifb_model_batch_logits.restype = vp
    L.ifb_model_batch_next_tokens.argtypes = [vp]
    L.ifb_model_batch_next_tokens.restype = vp
    L.ifb_model_hidden_ptr.argtypes = [vp, i32]
    L.ifb_model_hidden_ptr.restype = vp
    L.ifb_device_groups_parse.argtypes = [C.c_char_p, vp, i32, vp, vp, vp]
    L.ifb_split_layers.argtypes = [i32, i32, i32, vp, vp, vp]
    L.ifb_tensor_partition_type.argtypes = [C.c_char_p]
    L.ifb_tensor_device_dtype.argtypes = [i64, i32, i32, i64]
    L.ifb_rank_plan.argtypes = [C.c_char_p, i32, i32, vp]
    L.ifb_engine_resolve.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, vp, i32, vp]
    L.ifb_engine_create.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
    L.ifb_engine_destroy.argtypes = [vp]
    L.ifb_engine_destroy.restype = None
    L.ifb_engine_vocab_size.argtypes = [vp]
    L.ifb_engine_add_query.argtypes = [vp, vp, i32]
    L.ifb_engine_infer.argtypes = [vp, vp, vp]
    L.ifb_engine_commit.argtypes = [vp, i32, i32]
    L.ifb_engine_remove_query.argtypes = [vp, i32]
    L.ifb_llama2c_probe.argtypes = [C.c_char_p, vp, vp]
    L.ifb_model_load_llama2c.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, vp]
    L.ifb_model_load_llama2c_stage.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, i32, i32, vp]
    L.ifb_model_tp_export.argtypes = [vp, vp]
    L.ifb_model_tp_import.argtypes = [vp, vp, i32]
    return L


lib = _load()

# every symbol include/inferflow_b200.h declares (checked by tests/test_capi_symbols.py)
DECLARED = [
    "ifb_version", "ifb_last_error", "ifb_block_capacity", "ifb_block_bytes", "ifb_row_bytes",
    "ifb_dtype_from_name", "ifb_quantize_rows", "ifb_dequantize_rows", "ifb_gemv", "ifb_gemv_fused", "ifb_gemm_q",
    "ifb_rmsnorm", "ifb_rope", "ifb_silu_mul", "ifb_add", "ifb_argmax", "ifb_kv_cache_bytes",
    "ifb_kv_append", "ifb_kv_get_rows", "ifb_attn_workspace_bytes", "ifb_attn_decode", "ifb_attn_decode_append", "ifb_moe_topk", "ifb_moe_route", "ifb_moe_combine",
    "ifb_engine_resolve", "ifb_engine_create", "ifb_engine_destroy", "ifb_engine_vocab_size", "ifb_engine_add_query",
    "ifb_engine_infer", "ifb_engine_commit", "ifb_engine_remove_query", "ifb_llama2c_probe", "ifb_model_load_llama2c", "ifb_model_load_llama2c_stage", "ifb_model_create", "ifb_model_destroy", "ifb_model_weight_bytes", "ifb_model_kv_bytes_per_token",
    "ifb_model_init_synthetic", "ifb_model_load_tensor", "ifb_model_kv_reset", "ifb_model_decode",
    "ifb_model_logits", "ifb_model_next_token", "ifb_model_kernel_launches_per_step",
    "ifb_model_tp_export", "ifb_model_tp_import", "ifb_model_decode_next", "ifb_model_tensor_ptr",
    "ifb_memcpy_d2h", "ifb_memcpy_h2d", "ifb_memcpy_d2d", "ifb_device_synchronize", "ifb_model_issue_parts", "ifb_memcpy_d2h_async", "ifb_memcpy_h2d_async", "ifb_model_step_status", "ifb_model_prefill", "ifb_model_prefill_supported", "ifb_attn_prefill",
    "ifb_model_storage_dtype", "ifb_model_hidden_ptr", "ifb_step_plan", "ifb_model_batch_supported", "ifb_model_decode_batch", "ifb_model_batch_logits", "ifb_model_batch_next_tokens", "ifb_device_groups_parse", "ifb_split_layers", "ifb_tensor_partition_type",
    "ifb_tensor_device_dtype", "ifb_rank_plan",
]


def check(rc, what=""):
    if rc != 0:
        raise IfbError(f"{what} failed ({rc}): {lib.ifb_last_error().decode()}")


def dtype_from_name(name):
    return lib.ifb_dtype_from_name(name.encode())


def row_bytes(dt, cx):
    return lib.ifb_row_bytes(dt, cx)
