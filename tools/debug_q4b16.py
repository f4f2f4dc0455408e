import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import orc, inferflow_b200 as ib
dt = orc.Q4_B16
rng = np.random.default_rng(10 + dt)
for sigma, rows, cx in ((0.02, 64, 1024), (1.0, 33, 256), (1e-4, 8, 128)):
    src = orc.f16(rng.normal(0, sigma, (rows, cx))); src[0, :64] = 0.25
    src = orc.f16(np.clip(src, -0.99, 1.4))
    want = orc.oracle().quantize(dt, src)
    x = torch.from_numpy(src.astype(np.float16)).cuda()
    out = torch.zeros(want.size, dtype=torch.uint8, device='cuda')
    ib.check(ib.lib.ifb_quantize_rows(dt, C.c_void_p(out.data_ptr()), C.c_void_p(x.data_ptr()), rows, cx, None), 'q')
    got = out.cpu().numpy().reshape(want.shape)
    bad = np.argwhere(got != want)
    print(sigma, len(bad))
    for r, c in bad[:6]:
        blk = c // 10
        print(' row', r, 'blk', blk, 'byte', c % 10, 'got', got[r, blk*10:blk*10+10], 'want', want[r, blk*10:blk*10+10], 'src', src[r, blk*16:blk*16+16].tolist())
