"""Golden outputs of the REFERENCE ENGINE itself (CPU build of /root/reference, oracle/_ref/libifref_cpu.so,
driven through InferenceEngine::LoadConfig/Init/AddQuery/Infer/CommitInferenceResult) on a synthetic
stories110M-shaped llama (BASELINE configs[0] shape: L=12, d=768, f=2048, 12 heads, llama2.c RoPE order),
vocabulary cut to 8192 to keep the scratch model file small.

  python tests/golden/make_golden_engine.py        # needs `make -C oracle cpu` (i.e. /root/reference) first

Writes tests/golden/engine_stories110m.npz: the prompt, the greedy tokens, per step the reference's top-16
logits (ids + values) and the logits at 64 fixed probe ids."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_engine as R  # noqa: E402

HP = dict(dim=768, hidden=2048, layers=12, heads=12, kv_heads=12, vocab=8192, seq=256)
SEED = 1234
PROMPT = [1, 300, 5000, 77, 4242]
STEPS = 24
PROBE = np.arange(64) * 127 % HP["vocab"]


def run_reference(threads=8):
    w = R.synthetic_llama(HP, SEED)
    root = tempfile.mkdtemp(prefix="ifref_")
    try:
        ini = R.write_tree(root, HP, w, threads=threads)
        eng = R.RefEngine(ini)
        toks, logits, secs = eng.generate(PROMPT, STEPS, HP["vocab"])
        eng.close()
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return toks, logits, secs


if __name__ == "__main__":
    toks, logits, secs = run_reference()
    top_ids = np.argsort(-logits, axis=1)[:, :16].astype(np.int32)
    top_val = np.take_along_axis(logits, top_ids, axis=1).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "engine_stories110m.npz"), prompt=np.array(PROMPT, np.int32),
                        tokens=toks.astype(np.int32), top_ids=top_ids, top_val=top_val,
                        probe_ids=PROBE.astype(np.int32), probe_val=logits[:, PROBE].astype(np.float32),
                        seed=SEED, hp=np.array([HP[k] for k in ("dim", "hidden", "layers", "heads", "kv_heads", "vocab", "seq")]))
    print("tokens", toks.tolist(), "decode s", secs, "logit range", float(np.abs(logits).max()))
