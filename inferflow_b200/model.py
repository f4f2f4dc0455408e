"""Thin handle on the C++ decode runtime (ifb_model_* in include/inferflow_b200.h)."""
import ctypes as C
from dataclasses import dataclass

from .capi import DT, ModelSpecC, check, lib


@dataclass
class ModelSpec:
    layers: int
    d_model: int
    d_ffn: int
    heads: int
    kv_heads: int
    vocab: int
    max_ctx: int = 2048
    weight_dtype: int = DT.Q4_B32T1A
    kv_dtype: int = DT.Q8_B32T2
    lm_head_dtype: int = DT.F16
    rope_order: int = 2
    rope_theta: float = 10000.0
    norm_eps: float = 1e-5
    kq_scale: float = 1.0
    tp_rank: int = 0
    tp_size: int = 1
    experts: int = 0
    experts_top_k: int = 0
    moe_norm_top_k_prob: int = 1
    weight_repack: int = 0      # 1: Q3H_B64T1 matrices stored as Q4_B64T1 blocks with the same codes (lossless, faster to decode)
    kv_slots: int = 0           # KV caches side by side (<= 16): the sequences decode_batch advances together
    layer_begin: int = 0        # pipeline stage [layer_begin, layer_end); 0, 0 = the whole model
    layer_end: int = 0

    def to_c(self):
        return ModelSpecC(self.layers, self.d_model, self.d_ffn, self.heads, self.kv_heads, self.vocab,
                          self.max_ctx, self.weight_dtype, self.kv_dtype, self.lm_head_dtype,
                          self.rope_order, self.rope_theta, self.norm_eps, self.kq_scale, self.tp_rank,
                          self.tp_size, self.experts, self.experts_top_k, self.moe_norm_top_k_prob,
                          self.weight_repack, self.kv_slots, self.layer_begin, self.layer_end)


# shapes of SURVEY.md section 8
LLAMA2_7B = dict(layers=32, d_model=4096, d_ffn=11008, heads=32, kv_heads=32, vocab=32000)
LLAMA2_70B = dict(layers=80, d_model=8192, d_ffn=28672, heads=64, kv_heads=8, vocab=32000)
STORIES_110M = dict(layers=12, d_model=768, d_ffn=2048, heads=12, kv_heads=12, vocab=32000)


def probe_llama2c(path):
    """(dim, hidden, layers, heads, kv_heads, vocab, seq), shared_classifier of a llama2.c model file"""
    hv = (C.c_int32 * 7)()
    shared = C.c_int32(0)
    check(lib.ifb_llama2c_probe(path.encode(), hv, C.byref(shared)), "ifb_llama2c_probe")
    return tuple(hv), bool(shared.value)


class Model:
    def __init__(self, spec: ModelSpec, _handle=None):
        self.spec = spec
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
            return
        c = spec.to_c()
        check(lib.ifb_model_create(C.byref(self._h), C.byref(c)), "ifb_model_create")

    @classmethod
    def from_llama2c(cls, path, weight_dtype=DT.F16, kv_dtype=DT.F16, max_ctx=0, stream=0):
        """the reference's llama2.c file format (model_reader.cc:3245-3427) -> a loaded runtime"""
        (dim, hidden, layers, heads, kv_heads, vocab, seq), _ = probe_llama2c(path)
        h = C.c_void_p()
        check(lib.ifb_model_load_llama2c(C.byref(h), path.encode(), weight_dtype, kv_dtype, max_ctx, -1, stream),
              "ifb_model_load_llama2c")
        spec = ModelSpec(layers=layers, d_model=dim, d_ffn=hidden, heads=heads, kv_heads=kv_heads, vocab=vocab,
                         max_ctx=max_ctx if max_ctx > 0 else seq, weight_dtype=weight_dtype, kv_dtype=kv_dtype,
                         lm_head_dtype=DT.F16, rope_order=0)
        return cls(spec, _handle=h)

    def close(self):
        if self._h:
            lib.ifb_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def weight_bytes(self):
        return lib.ifb_model_weight_bytes(self._h)

    def kv_bytes_per_token(self, ctx_len):
        return lib.ifb_model_kv_bytes_per_token(self._h, ctx_len)

    def init_synthetic(self, seed=1234, stream=0):
        check(lib.ifb_model_init_synthetic(self._h, seed, stream), "ifb_model_init_synthetic")

    def load_tensor(self, name, array, stream=0):
        import numpy as np
        a = np.ascontiguousarray(array)
        dt = {np.dtype(np.float32): DT.F32, np.dtype(np.float16): DT.F16}[a.dtype]
        rows, cols = (a.shape if a.ndim == 2 else (1, a.shape[0]))
        check(lib.ifb_model_load_tensor(self._h, name.encode(), a.ctypes.data, dt, rows, cols, stream),
              f"ifb_model_load_tensor({name})")

    def tensor_ptr(self, name):
        p = C.c_void_p()
        n = C.c_int64()
        check(lib.ifb_model_tensor_ptr(self._h, name.encode(), C.byref(p), C.byref(n)), "ifb_model_tensor_ptr")
        return p.value, n.value

    @property
    def storage_weight_dtype(self):
        """element type the layer matrices are stored in (differs from spec.weight_dtype under weight_repack)"""
        return int(lib.ifb_model_storage_dtype(self._h, 0))

    @property
    def hidden_in_ptr(self):
        """device buffer a pipeline stage with layer_begin > 0 reads its input hidden state from"""
        return lib.ifb_model_hidden_ptr(self._h, 0)

    @property
    def hidden_out_ptr(self):
        """device buffer a pipeline stage leaves its output hidden state in"""
        return lib.ifb_model_hidden_ptr(self._h, 1)

    def read_tensor(self, name, nbytes=None):
        """device bytes of a named tensor as a numpy uint8 array (blocking)"""
        import numpy as np
        p, n = self.tensor_ptr(name)
        n = n if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        check(lib.ifb_memcpy_d2h(out.ctypes.data, p, n), "ifb_memcpy_d2h")
        return out

    def kv_reset(self, ctx_len=0, fill_synthetic=False, seed=99, stream=0):
        check(lib.ifb_model_kv_reset(self._h, ctx_len, 1 if fill_synthetic else 0, seed, stream), "ifb_model_kv_reset")

    def decode(self, token_id, pos, stream=0, token_dev=None):
        check(lib.ifb_model_decode(self._h, token_id, token_dev, pos, stream), "ifb_model_decode")

    def prefill(self, tokens, pos0=0, stream=0):
        """the prompt in one pass (ifb_model_prefill); logits / next token of the last prompt token as after decode()"""
        import numpy as np
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        check(lib.ifb_model_prefill(self._h, t.ctypes.data, len(t), pos0, stream), "ifb_model_prefill")

    def decode_batch(self, tokens, slots, positions, stream=0):
        """one token for each of n sequences (KV slot, position); returns (logits fp16 [n][vocab], next tokens [n])"""
        import numpy as np
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        ps = np.ascontiguousarray(positions, dtype=np.int32)
        n = len(t)
        check(lib.ifb_model_decode_batch(self._h, t.ctypes.data, sl.ctypes.data, ps.ctypes.data, n, stream), "ifb_model_decode_batch")
        check(lib.ifb_device_synchronize(), "sync")
        logits = np.empty((n, self.spec.vocab), dtype=np.float16)
        check(lib.ifb_memcpy_d2h(logits.ctypes.data, lib.ifb_model_batch_logits(self._h), logits.nbytes), "d2h")
        nxt = np.empty(n, dtype=np.int32)
        check(lib.ifb_memcpy_d2h(nxt.ctypes.data, lib.ifb_model_batch_next_tokens(self._h), 4 * n), "d2h")
        return logits, nxt

    @property
    def batch_supported(self):
        return bool(lib.ifb_model_batch_supported(self._h))

    @property
    def prefill_supported(self):
        return bool(lib.ifb_model_prefill_supported(self._h))

    def decode_next(self, stream=0):
        check(lib.ifb_model_decode_next(self._h, stream), "ifb_model_decode_next")

    # -- tensor parallelism ------------------------------------------------------------------------------
    def tp_export(self):
        """64-byte IPC handle of this rank's exchange region (partial sums, logits, flags)."""
        buf = (C.c_uint8 * 64)()
        check(lib.ifb_model_tp_export(self._h, buf), "ifb_model_tp_export")
        return bytes(buf)

    def tp_import(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        check(lib.ifb_model_tp_import(self._h, blob, len(handles)), "ifb_model_tp_import")

    def tp_connect(self, dist, group=None):
        """Exchange the IPC handles over the process group (control plane only; the data path is the kernels'
        own peer-memory stores) and map every peer's exchange region."""
        if self.spec.tp_size == 1:
            return
        handles = exchange_handles(dist, self.tp_export(), group)
        assert len(handles) == self.spec.tp_size
        self.tp_import(handles)
        dist.barrier(group)

    def issue_parts(self, parts, stream=0):
        n = lib.ifb_model_issue_parts(self._h, parts, stream)
        if n < 0:
            check(n, "ifb_model_issue_parts")
        return n

    @property
    def logits_ptr(self):
        return lib.ifb_model_logits(self._h)

    @property
    def next_token_ptr(self):
        return lib.ifb_model_next_token(self._h)

    @property
    def launches_per_step(self):
        return lib.ifb_model_kernel_launches_per_step(self._h)

    def step_status(self):
        """(uses the persistent step kernel?, error record {code, op, detail, cta} of a step that missed a deadline)"""
        rec = (C.c_int32 * 4)()
        on = lib.ifb_model_step_status(self._h, rec)
        return bool(on), tuple(rec)


def exchange_handles(dist, handle: bytes, group=None):
    """all_gather of fixed-size opaque handles; works over nccl (device tensors) and gloo (host tensors)."""
    import torch
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    world = dist.get_world_size(group)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def shard_rows(rows, rank, size):
    """[begin, end) of a tensor's rows held by `rank` when the rows are split `size` ways
    (column-parallel weights wq/wk/wv/w1/w3 and the lm_head; reference: TensorPartitionType::BY_COL,
    network_builder.cc:1138-1215 splits ne[1])."""
    assert rows % size == 0, f"{rows} rows do not split {size} ways"
    n = rows // size
    return rank * n, (rank + 1) * n


def shard_cols(cols, rank, size, block=64):
    """[begin, end) of the input columns held by `rank` for row-parallel weights (wo, w2; reference BY_ROW);
    shard edges stay on quantization-block boundaries so every shard quantizes to the same block bytes."""
    assert cols % (size * block) == 0, f"{cols} columns do not split {size} ways on {block}-element blocks"
    n = cols // size
    return rank * n, (rank + 1) * n


def resolve_config(ini_path, section="transformer_engine", data_root_dir=""):
    """what Engine(ini_path) would load, without a GPU: (model file, weight dtype, kv dtype, max_context_len,
    qk_column_order, return_output_tensors)"""
    path = C.create_string_buffer(4096)
    info = (C.c_int32 * 5)()
    check(lib.ifb_engine_resolve(ini_path.encode(), section.encode(), data_root_dir.encode(), path, 4096, info),
          "ifb_engine_resolve")
    return (path.value.decode(),) + tuple(info)


class Engine:
    """InferenceEngine-shaped door (LoadConfig/Init, AddQuery, Infer, CommitInferenceResult, RemoveQuery) for one
    greedy query; reads the reference's .ini + model-spec json (see ifb_engine.cu)."""

    def __init__(self, ini_path, section="transformer_engine", data_root_dir=""):
        self._h = C.c_void_p()
        check(lib.ifb_engine_create(C.byref(self._h), ini_path.encode(), section.encode(), data_root_dir.encode()),
              "ifb_engine_create")
        self.vocab = lib.ifb_engine_vocab_size(self._h)

    def add_query(self, tokens):
        import numpy as np
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        qid = lib.ifb_engine_add_query(self._h, t.ctypes.data, len(t))
        if qid < 0:
            check(qid, "ifb_engine_add_query")
        return qid

    def infer(self, want_logits=False):
        import numpy as np
        nxt = C.c_int32(0)
        logits = np.zeros(self.vocab, dtype=np.float32) if want_logits else None
        check(lib.ifb_engine_infer(self._h, C.byref(nxt), logits.ctypes.data if want_logits else None), "ifb_engine_infer")
        return nxt.value, logits

    def commit(self, token, is_end=False):
        check(lib.ifb_engine_commit(self._h, int(token), 1 if is_end else 0), "ifb_engine_commit")

    def remove_query(self, qid):
        check(lib.ifb_engine_remove_query(self._h, int(qid)), "ifb_engine_remove_query")

    def generate(self, prompt, max_new, want_logits=False):
        """the loop of oracle/ref_cpu_shim.cc ifref_engine_generate, on this engine"""
        qid = self.add_query(prompt)
        assert qid > 0
        toks, logs = [], []
        for step in range(max_new):
            t, lg = self.infer(want_logits)
            toks.append(t)
            if want_logits:
                logs.append(lg)
            self.commit(t, is_end=(step + 1 >= max_new))
        return toks, logs

    def close(self):
        if self._h:
            lib.ifb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
