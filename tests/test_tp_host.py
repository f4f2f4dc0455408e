"""Host side of the tensor-parallel path on CPU: two processes over gloo.

What runs here is everything around the kernels: the shard arithmetic ifb_model_load_tensor applies (BY_COL /
BY_ROW of reference network_builder.cc:1138-1215), the handle exchange Model.tp_connect performs over the process
group, and the shard-sum order of the exchange (oracle restatement, tests/orc_model.py) against the unsharded
oracle -- with the partial sums actually travelling between the two processes."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import orc
        from inferflow_b200.model import exchange_handles, shard_cols, shard_rows
        o = orc.oracle()

        # 1. the handle exchange of tp_connect: every rank ends with every rank's 64 bytes, in rank order
        mine = bytes([rank * 16 + (i % 16) for i in range(64)])
        got = exchange_handles(dist, mine)
        assert len(got) == world and all(len(h) == 64 for h in got)
        for r in range(world):
            assert got[r] == bytes([r * 16 + (i % 16) for i in range(64)])

        # 2. a column-parallel then row-parallel pair (w1 -> w2 of one FFN) on sharded quantized weights: each
        #    rank quantizes ITS shard of the fp16 matrix; the bytes equal the slice of the unsharded quantization
        rng = np.random.default_rng(5)       # same stream on both ranks
        d, f, dt = 256, 512, orc.Q4_B32T1A
        w1 = (0.06 * rng.standard_normal((f, d))).astype(np.float16).astype(np.float32)
        w2 = (0.06 * rng.standard_normal((d, f))).astype(np.float16).astype(np.float32)
        x = rng.standard_normal(d).astype(np.float16).astype(np.float32)
        r0, r1 = shard_rows(f, rank, world)
        c0, c1 = shard_cols(f, rank, world)
        assert (r0, r1) == (c0, c1)          # the FFN's hidden slice is the same on both sides
        w1_full = o.quantize(dt, w1)
        w2_full = o.quantize(dt, w2)
        w1_sh = o.quantize(dt, np.ascontiguousarray(w1[r0:r1]))
        w2_sh = o.quantize(dt, np.ascontiguousarray(w2[:, c0:c1]))
        rb = orc.row_bytes(dt, f)
        assert np.array_equal(w1_sh.reshape(-1), w1_full.reshape(f, -1)[r0:r1].reshape(-1))
        assert np.array_equal(w2_sh.reshape(d, -1), w2_full.reshape(d, rb)[:, rb // world * rank: rb // world * (rank + 1)])

        # 3. the exchange: partial sums rounded to fp16, added in rank order on every rank -> identical result
        #    everywhere, and equal to the oracle's tp restatement computed from the UNSHARDED weights
        xq = o.quantize(orc.Q8_B32T2, x[None, :])
        h = orc.f16(o.gemv_ax8(dt, w1_sh.reshape(-1), xq, r1 - r0, d))                 # this rank's hidden slice
        part = orc.f16(o.gemv_ax8(dt, w2_sh.reshape(-1), o.quantize(orc.Q8_B32T2, h[None, :]), d, c1 - c0))
        parts = [torch.empty(d, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(part.astype(np.float32)))
        acc = parts[0].numpy()
        for p in parts[1:]:
            acc = orc.f16(acc + p.numpy())
        both = [torch.empty(d, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(both, torch.from_numpy(acc.astype(np.float32)))
        assert all(torch.equal(b, both[0]) for b in both)

        h_full = orc.f16(o.gemv_ax8(dt, w1_full.reshape(-1), xq, f, d))
        want = None
        for r in range(world):
            a0, a1 = shard_cols(f, r, world)
            sh = o.quantize(dt, np.ascontiguousarray(w2[:, a0:a1]))
            pr = orc.f16(o.gemv_ax8(dt, sh.reshape(-1), o.quantize(orc.Q8_B32T2, h_full[None, a0:a1]), d, a1 - a0))
            want = pr if want is None else orc.f16(want + pr)
        assert np.array_equal(acc, want)
        # and close to (not bit-equal with) the unsharded product: only the rounding order differs
        full = orc.f16(o.gemv_ax8(dt, w2_full.reshape(-1), o.quantize(orc.Q8_B32T2, h_full[None, :]), d, f))
        assert np.abs(full - acc).max() <= 2 ** -9 * np.abs(full).max() + 2 ** -10
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))


def test_two_rank_gloo_exchange_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_shard_helpers_reject_ragged_splits():
    from inferflow_b200.model import shard_cols, shard_rows
    assert shard_rows(32000, 3, 8) == (12000, 16000)
    assert shard_cols(11008, 1, 2) == (5504, 11008)
    with pytest.raises(AssertionError):
        shard_rows(10, 0, 3)
    with pytest.raises(AssertionError):
        shard_cols(11008, 0, 8, block=64)     # 1376 per rank is not a multiple of 64 -> refuse, do not mis-quantize
