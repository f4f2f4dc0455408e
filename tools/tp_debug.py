"""Stage-by-stage comparison of a tensor-parallel decode step with the single-GPU runtime (run under torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import inferflow_b200 as ib  # noqa: E402
import orc  # noqa: E402
from tp_worker import host_weights, load_all  # noqa: E402


def rd(m, name, n):
    ib.check(ib.lib.ifb_device_synchronize(), "sync")
    return m.read_tensor(name, n * 2).view(np.float16).astype(np.float32)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    layers = int(os.environ.get("DBG_LAYERS", "1"))
    base = dict(layers=layers, vocab=1024, max_ctx=128, weight_dtype=orc.Q4_B32T1A, kv_dtype=orc.Q8_B32T2,
                heads=8, kv_heads=8, d_model=512, d_ffn=1024)
    spec = ib.ModelSpec(tp_rank=rank, tp_size=world, **base)
    w = host_weights(spec, 100)
    m = ib.Model(spec); load_all(m, w); m.tp_connect(dist)
    ref = ib.Model(ib.ModelSpec(**base)); load_all(ref, w)
    d, f = spec.d_model, spec.d_ffn
    qd = d // world; fl = f // world
    for step in range(2):
        m.decode(5, step); ref.decode(5, step)
        q_t, q_r = rd(m, "qkv", 3 * qd), rd(ref, "qkv", 3 * d)
        want = np.concatenate([q_r[i * d + rank * qd: i * d + (rank + 1) * qd] for i in range(3)])
        out = [f"rank {rank} step {step}:", f"qkv {np.abs(q_t - want).max():.3e}"]
        out.append(f"attn {np.abs(rd(m, 'attn_out', qd) - rd(ref, 'attn_out', d)[rank * qd:(rank + 1) * qd]).max():.3e}")
        out.append(f"x1 {np.abs(rd(m, 'x1', d) - rd(ref, 'x1', d)).max():.3e}")
        out.append(f"hbuf {np.abs(rd(m, 'hbuf', fl) - rd(ref, 'hbuf', f)[rank * fl:(rank + 1) * fl]).max():.3e}")
        out.append(f"x0 {np.abs(rd(m, 'x', d) - rd(ref, 'x', d)).max():.3e} (range {np.abs(rd(ref, 'x', d)).max():.2f})")
        out.append(f"logits {np.abs(rd(m, 'logits', spec.vocab) - rd(ref, 'logits', spec.vocab)).max():.3e}")
        print("  ".join(out), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
