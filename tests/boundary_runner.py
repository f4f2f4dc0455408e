"""Run the reference's CUDA ENGINE (one of oracle/_ref/libifref_engine_{stock,ifb}.so) on a config tree and save the
greedy tokens + logits. Own process per library: both export the same C++ symbols.
usage: boundary_runner.py <lib.so> <ini> <out.npz> <vocab> <max_new> <prompt ids...>"""
import ctypes as C
import os
import sys

import numpy as np

lib_path, ini, out, vocab, max_new = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
prompt = np.array([int(x) for x in sys.argv[6:]], dtype=np.int32)
lib = C.CDLL(lib_path, mode=C.RTLD_GLOBAL)
lib.ifref_engine_create.restype = C.c_void_p
lib.ifref_engine_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
lib.ifref_engine_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
lib.ifref_engine_destroy.argtypes = [C.c_void_p]
data_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(ini))), "data", "inferflow") + "/"
h = lib.ifref_engine_create(ini.encode(), b"transformer_engine", data_root.encode())
if not h:
    sys.exit("the reference engine failed to load " + ini)
toks = np.zeros(max_new, dtype=np.int32)
logits = np.zeros((max_new, vocab), dtype=np.float32)
secs = C.c_double(0.0)
n = lib.ifref_engine_generate(h, prompt.ctypes.data, len(prompt), max_new, toks.ctypes.data, logits.ctypes.data, vocab, C.byref(secs))
if n < 0:
    sys.exit(f"generate failed ({n})")
np.savez(out, tokens=toks[:n], logits=logits[:n], decode_seconds=secs.value)
lib.ifref_engine_destroy(h)
