"""Drive the UNMODIFIED reference engine (CPU build, oracle/_ref/libifref_cpu.so) on a synthetic model.

TEST / BENCH INFRASTRUCTURE. Writes, into a scratch directory, everything InferenceEngine::LoadConfig + Init read
(reference inference_engine.h:39,118): an .ini in the layout of bin/llm_inference.tiny.ini, a model
specification json in the layout of data/models/llama2.c/model_spec_15m.json, a llama2.c-format weight file
(7 x int32 header, fp32 tensors in the order of model_reader.cc:3339-3427) and a llama2.c tokenizer.bin
(ReadVocabulary_Format2, model_reader.cc:1361-1405). Nothing is copied from the reference tree:
all four files are generated from the numbers passed in."""
import ctypes as C
import json
import os
import struct

import numpy as np

import orc


def write_tokenizer(path, vocab):
    """llama2.c tokenizer.bin as ReadVocabulary_Format2 parses it (model_reader.cc:1361-1405):
    uint32 max_token_len, then per token: float score, uint32 len, bytes"""
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 16))
        for i in range(vocab):
            s = ("<unk>", "<s>", "</s>")[i].encode() if i < 3 else ("<%d>" % i).encode()
            f.write(struct.pack("<fI", -float(i), len(s)) + s)


def synthetic_llama(hp, seed=1234):
    """seeded fp32 weights whose values are fp16-representable (so a loader that stores fp16 holds the same
    numbers): embeddings N(0,1), matrices N(0,0.04), norms 1 + 0.1 N(0,1), separate classifier"""
    rng = np.random.default_rng(seed)
    d, f, L, V = hp["dim"], hp["hidden"], hp["layers"], hp["vocab"]
    hd = d // hp["heads"]
    kvd = hp["kv_heads"] * hd

    def r(shape, std, mean=0.0):
        return (mean + std * rng.standard_normal(shape)).astype(np.float16).astype(np.float32)
    w = {"tok_embeddings": r((V, d), 1.0), "output_norm": r((d,), 0.1, 1.0), "output": r((V, d), 0.04)}
    for l in range(L):
        p = f"layers.{l}."
        w[p + "attention_norm"] = r((d,), 0.1, 1.0)
        w[p + "ffn_norm"] = r((d,), 0.1, 1.0)
        w[p + "wq"] = r((d, d), 0.04); w[p + "wk"] = r((kvd, d), 0.04); w[p + "wv"] = r((kvd, d), 0.04)
        w[p + "wo"] = r((d, d), 0.04)
        w[p + "w1"] = r((f, d), 0.04); w[p + "w3"] = r((f, d), 0.04); w[p + "w2"] = r((d, f), 0.04)
    return w


def write_llama2c(path, hp, w):
    """hp: dict(dim, hidden, layers, heads, kv_heads, vocab, seq); w: name -> fp32 arrays (full, unsharded):
    tok_embeddings [V,d], layers.N.{attention_norm [d], wq, wk, wv, wo, ffn_norm, w1, w2, w3}, output_norm [d]"""
    L, d, hd = hp["layers"], hp["dim"], hp["dim"] // hp["heads"]
    with open(path, "wb") as f:
        shared = "output" not in w       # llama2.c: a negative vocab size announces a separate classifier
        f.write(struct.pack("<7i", d, hp["hidden"], L, hp["heads"], hp["kv_heads"],
                            hp["vocab"] if shared else -hp["vocab"], hp["seq"]))

        def put(a):
            a = np.ascontiguousarray(a, dtype=np.float32)
            f.write(memoryview(a.reshape(-1)))
        put(w["tok_embeddings"])
        for name in ("attention_norm", "wq", "wk", "wv", "wo", "ffn_norm", "w1", "w2", "w3"):
            for l in range(L):
                put(w[f"layers.{l}.{name}"])
        put(w["output_norm"])
        # llama2.c stores freq_cis_real/imag tables here (2 x seq x head_dim/2 floats); the reference reader skips
        # exactly training_context_len * head_size BYTES (model_reader.cc:3420-3421) before an unshared classifier,
        # so that is what has to be there for it to find the classifier
        f.write(bytes(hp["seq"] * hd))
        if not shared:
            put(w["output"])


def write_tree(root, hp, w, threads, name="synth", return_logits=True):
    """<root>/bin/ref.ini, <root>/data/inferflow/, <root>/data/models/<name>/{model.bin,tokenizer.bin,model_spec.json}"""
    mdir = os.path.join(root, "data", "models", name)
    os.makedirs(mdir, exist_ok=True)
    os.makedirs(os.path.join(root, "data", "inferflow", "logs"), exist_ok=True)
    os.makedirs(os.path.join(root, "bin"), exist_ok=True)
    write_tokenizer(os.path.join(mdir, "tokenizer.bin"), hp["vocab"])
    write_llama2c(os.path.join(mdir, "model.bin"), hp, w)
    spec = {"config_file": "", "model_files": ["model.bin"], "model_file_format": "llama2.c",
            "tokenizer_file": "tokenizer.bin", "tokenization_algorithm": "bpe", "generation_config": "",
            "qkv_format": 1,
            "network_structure": {"type": "transformer.llama", "normalization_function": "rms",
                                  "activation_function": "silu", "position_embedding": "rope",
                                  "qk_column_order": 0, "tensor_name_prefix": "", "tensor_name_mapping": {}}}
    with open(os.path.join(mdir, "model_spec.json"), "w") as f:
        json.dump(spec, f, indent=1)
    return write_ini(root, hp, threads, name, return_logits)


def write_ini(root, hp, threads, name="synth", return_logits=True):
    """(re)write <root>/bin/ref.ini only -- e.g. another cpu_threads value for the same model files"""
    ini = f"""[transformer_engine]
models = {name}
devices = 0
decoder_cpu_layer_count = {hp['layers']}
cpu_threads = {threads}
max_concurrent_queries = 1
return_output_tensors = {'true' if return_logits else 'false'}
is_study_mode = false
show_tensors = false

[model.{name}]
model_dir = ${{data_root_dir}}../models/{name}/
model_specification_file = model_spec.json
device_weight_data_type = F16
device_kv_cache_data_type = F16
host_kv_cache_percent = 0
max_context_len = {hp['seq']}
prompt_template = {{bos}}{{query}}

[prompt_templates]
prompt_template_count = 0

[app_env.base]
data_root_dir = ${{config_dir}}../data/inferflow/
require_enter_key_to_exit = false

[app_env.logging]
enable_logging = 0
log_dir = ${{data_root_dir}}logs
log_name = ref_engine

[app_env.status_manager]
enable_monitoring = 0
status_file = ${{data_root_dir}}/ref_engine.status
listening_port = 9877
"""
    p = os.path.join(root, "bin", "ref.ini")
    with open(p, "w") as f:
        f.write(ini)
    return p


class RefEngine:
    def __init__(self, ini_path):
        self.ref = orc.RefCpu()
        # the tools pass data_root_dir from [app_env.base] (app_environment.cc); LoadConfig takes it as a macro
        data_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(ini_path))), "data", "inferflow") + "/"
        self.h = self.ref.lib.ifref_engine_create(ini_path.encode(), b"transformer_engine", data_root.encode())
        if not self.h:
            raise RuntimeError("the reference engine failed to load " + ini_path)

    def generate(self, prompt, max_new, vocab, want_logits=True):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        toks = np.zeros(max_new, dtype=np.int32)
        logits = np.zeros((max_new, vocab), dtype=np.float32) if want_logits else None
        secs = C.c_double(0.0)
        n = self.ref.lib.ifref_engine_generate(self.h, prompt.ctypes.data, len(prompt), max_new, toks.ctypes.data,
                                               logits.ctypes.data if want_logits else None, vocab, C.byref(secs))
        if n < 0:
            raise RuntimeError(f"reference engine: generate failed ({n})")
        return toks[:n], (logits[:n] if want_logits else None), secs.value

    def close(self):
        if self.h:
            self.ref.lib.ifref_engine_destroy(self.h)
            self.h = None
