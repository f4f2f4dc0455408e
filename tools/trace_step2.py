"""Timeline of the persistent step kernel (IFB_TRACE=1): per op and CTA four globaltimer stamps --
0 staging entered (all rows of the previous op stored), 1 dependency met (+ residual updated), 2 input staged,
3 first unit of the op in hand. Printed for one layer in the middle of the model and for the lm_head, us relative to
the layer's first stamp: min / median / max over the CTAs."""
import os, sys
os.environ["IFB_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import inferflow_b200 as ib
torch.zeros(1, device="cuda")
spec = ib.ModelSpec(max_ctx=2048, **ib.LLAMA2_7B)
m = ib.Model(spec); m.init_synthetic(1234); m.kv_reset(2047, True)
for _ in range(4):
    m.decode(1, 2047)
ib.check(ib.lib.ifb_device_synchronize(), "sync")
on, rec = m.step_status()
print("step kernel:", on, rec)
n_ops = 5 * spec.layers + 1
grid = 148
raw = m.read_tensor("trace").view(np.uint64)[: n_ops * grid * 8].reshape(n_ops, grid, 8).astype(np.int64)
names = ["QKV", "ATT", "WO", "GLU", "W2"]
t_first = raw[0, :, 0].min()
t_last = raw[n_ops - 1, :, 3].max()
print(f"first stamp -> first lm_head unit: {(t_last - t_first) / 1000.0:.1f} us")
for L in (1, 16):
    t0 = raw[5 * L, :, 0].min()
    print(f"--- layer {L} (us from its first stamp); next layer starts at {(raw[5 * L + 5, :, 0].min() - t0) / 1000.0:.2f}")
    for k in range(5):
        r = (raw[5 * L + k] - t0) / 1000.0
        def f(c):
            v = r[:, c]
            return f"{v.min():7.2f}/{np.median(v):7.2f}/{v.max():7.2f}"
        print(f"{names[k]:4s} enter {f(0)}  dep {f(1)}  staged {f(2)}  first_unit {f(3)}")
        c = raw[5 * L + k]
        mhz = 1965.0
        print(f"       warp0 slot walk med {np.median(c[:, 5]) / mhz:6.2f} us | producer: wait-for-free-slot med {np.median(c[:, 6]) / mhz:6.2f} us  units/CTA min {c[:, 7].min()} med {int(np.median(c[:, 7]))} max {c[:, 7].max()}")
r = (raw[n_ops - 1] - raw[n_ops - 1, :, 0].min()) / 1000.0
print("HEAD enter %.2f/%.2f dep %.2f staged %.2f first_unit %.2f" % (r[:, 0].min(), r[:, 0].max(), np.median(r[:, 1]), np.median(r[:, 2]), np.median(r[:, 3])))
# per-layer period
starts = np.array([raw[5 * l, :, 0].min() for l in range(spec.layers)])
print("layer period us: median %.2f min %.2f max %.2f" % (np.median(np.diff(starts)) / 1000, np.diff(starts).min() / 1000, np.diff(starts).max() / 1000))
