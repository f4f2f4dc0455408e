"""Timeline probe of one decode step (IFB_TRACE=1): per kernel launch, ns relative to the first kernel.
slots: 0 kernel entry, 1 after griddepcontrol.wait, 2 activation staged, 3 first tile landed,
4 consumer done, 5 first TMA issued, 6 last TMA issued."""
import os, sys
os.environ["IFB_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import inferflow_b200 as ib
graph = "--graph" in sys.argv
if not graph:
    os.environ["IFB_NO_GRAPH"] = "1"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
torch.zeros(1, device="cuda")
spec = ib.ModelSpec(max_ctx=2048, tp_rank=rank, tp_size=world, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(2047, True)
if world > 1:   # torchrun: one process per GPU, the timeline printed is rank 0's
    import torch.distributed as dist
    dist.init_process_group("nccl")
    m.tp_connect(dist)
for _ in range(4):
    m.decode(1, 2047)
ib.check(ib.lib.ifb_device_synchronize(), "sync")
if world > 1:
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group(); sys.exit(0)
raw = m.read_tensor("trace").view(np.uint64).reshape(512, 160, 8).astype(np.int64)
n = m.launches_per_step
t00 = raw[1][:148, 0][raw[1][:148, 0] > 0].min()
names = {0: "embed"}
print("launches", n)
rows = []
for k in range(1, min(n, 16)):
    t = raw[k]
    act = t[:, 0] > 0
    if not act.any():
        continue
    def st(col, f):
        v = t[act, col]; v = v[v > 0]
        return (f(v) - t00) / 1000.0 if len(v) else float("nan")
    if "--attn" in sys.argv and act.sum() < 148:
        print(f"k{k:3d} attn: wait {st(1,np.median):7.2f} tile {st(3,np.median):7.2f} qk {st(5,np.median):7.2f} softmax {st(6,np.median):7.2f} pv_done[med {st(4,np.median):7.2f} max {st(4,np.max):7.2f}] partial_written[med {st(7,np.median):7.2f} max {st(7,np.max):7.2f}]")
        continue
    print(f"k{k:3d} ctas {act.sum():3d} entry[min {st(0,np.min):8.2f} max {st(0,np.max):8.2f}] wait_done[min {st(1,np.min):8.2f} max {st(1,np.max):8.2f}] staged[max {st(2,np.max):8.2f}] "
          f"first_tile[min {st(3,np.min):8.2f} max {st(3,np.max):8.2f}] done[min {st(4,np.min):8.2f} max {st(4,np.max):8.2f}] staged[med {st(2,np.median):8.2f}] cta_done[med {st(5,np.median):8.2f} max {st(5,np.max):8.2f}] peers_seen[min {st(6,np.min):8.2f} med {st(6,np.median):8.2f} max {st(6,np.max):8.2f}] published[{st(7,np.max):8.2f}]")
