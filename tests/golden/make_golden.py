"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libifref_cpu.so,
built by oracle/Makefile from /root/reference). Run in the authoring container only:

    make -C oracle cpu && python tests/golden/make_golden.py

The reference ships no golden vectors; these fixtures are the pin:
  codec_<TYPE>.npz : src (fp16-representable fp32) -> reference blocks, reference
                     dequantization, reference GetInt4 words (dp4a formats)
  moe_topk.npz     : router scores -> HostTensorOpr::BuildRowsForMoE ids / weights
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402


def codec_inputs(rng, cap):
    rows = []
    n = 4 * 64  # 4 B64 blocks / 8 B32 blocks per row
    rows.append(rng.normal(0, 0.02, n))                    # weight-like
    rows.append(rng.normal(0, 1.0, n))                     # activation-like
    rows.append(rng.uniform(-1.2, 1.2, n))                 # inside Q4_B16's base range
    rows.append(np.full(n, 0.125))                         # constant block -> scale < 1e-5
    rows.append(np.zeros(n))
    r = rng.normal(0, 0.5, n); r[::7] = 3.0; r[3::11] = -2.5   # outliers
    rows.append(r)
    rows.append(0.05 * np.sin(0.37 * np.arange(n)))        # the survey's probe signal
    rows.append(np.linspace(-1, 1, n))                     # exact ties on the rounding grid
    rows.append(rng.normal(0, 1e-4, n))                    # tiny range around the 1e-5 threshold
    x = np.stack(rows).astype(np.float32)
    return orc.f16(x)


def main():
    ref = orc.RefCpu()
    rng = np.random.default_rng(1234)
    for dt in orc.ALL_QUANT:
        src = codec_inputs(rng, orc.CAP[dt])
        if dt == orc.Q4_B16:
            # Q4_B16 stores base as a byte = base*100+100: valid for base in [-1, 1.55] only (outside it the
            # reference's float->uint8 conversion is undefined behaviour); keep the fixture inside the domain
            src = orc.f16(np.clip(src, -0.99, 1.4))
        blocks = ref.quantize(dt, src)
        deq = ref.dequantize(dt, blocks, src.shape[1])
        out = dict(src=src, blocks=blocks, deq=deq)
        if dt in orc.AX8_TYPES:
            out["int4"] = ref.get_int4(dt, blocks, src.shape[1])
        np.savez_compressed(os.path.join(HERE, f"codec_{orc.NAMES[dt]}.npz"), **out)
        print(orc.NAMES[dt], blocks.shape, "ref sizeof", ref.lib.ifref_sizeof_block(dt))

    scores = rng.random((64, 8)).astype(np.float32)
    scores = scores / scores.sum(1, keepdims=True)
    scores[3, 2] = scores[3, 5]          # tie: first wins
    scores[4, :] = 0.0; scores[4, 1] = 1.0   # < 1e-5 dropped
    scores[5, :] = 0.125                  # all equal
    s16 = scores.astype(np.float16)
    outs = {}
    for k in (1, 2, 4):
        for norm in (0, 1):
            ids, wts, cnt = ref.moe_topk(s16, k, bool(norm))
            outs[f"ids_k{k}_n{norm}"] = ids
            outs[f"w_k{k}_n{norm}"] = wts
            outs[f"cnt_k{k}_n{norm}"] = cnt
    np.savez_compressed(os.path.join(HERE, "moe_topk.npz"), scores=s16, **outs)
    print("moe_topk ok")


if __name__ == "__main__":
    main()
