"""Host-side oracle of one llama decode step (TEST INFRASTRUCTURE): the layer schedule of
GpuInferenceWorker::ProcessGpuLayer (inference_worker.cc:762-977) restated with the oracle's
operators, every intermediate tensor rounded to fp16 where the reference stores half."""
import numpy as np

import orc

AX8 = set(orc.AX8_TYPES)


class OracleModel:
    def __init__(self, spec, tensors, tp=1):
        """tensors: name -> raw uint8 bytes as stored on the device (reference block layout), UNSHARDED.
        tp > 1: the row-parallel products (wo, w2) are computed as tp partial sums over column shards, each
        stored as half and added in rank order, the residual last -- the order the reference's tensor-parallel
        merge uses (MergeTensors: inference_worker.cc:2264-2289, then the layer's residual Add)."""
        self.tp = tp
        self.s = spec
        self.t = tensors
        self.hd = spec.d_model // spec.heads
        self.K = [[] for _ in range(spec.layers)]
        self.V = [[] for _ in range(spec.layers)]
        self.o = orc.oracle()

    def _gemv(self, name, dt, x, rows, cols):
        w = self.t[name][: orc.row_bytes(dt, cols) * rows] if dt != orc.F16 else self.t[name][: rows * cols * 2]
        if dt == orc.F16:
            return self.o.gemv_f16(w.view(np.float16).reshape(rows, cols), x, rows, cols)
        if dt in AX8:
            xq = self.o.quantize(orc.Q8_B32T2, x[None, :])
            return self.o.gemv_ax8(dt, w, xq, rows, cols)
        return self.o.gemv_half(dt, w, x, rows, cols)

    def _gemv_rowpar(self, name, dt, x, rows, cols, resid):
        """round_h(W x) + resid, or its tensor-parallel restatement"""
        if self.tp == 1:
            return orc.f16(orc.f16(self._gemv(name, dt, x, rows, cols)) + resid)
        rb = orc.row_bytes(dt, cols) if dt != orc.F16 else cols * 2
        w = self.t[name][: rb * rows].reshape(rows, rb)
        nc = cols // self.tp
        sb = rb // self.tp
        acc = None
        for r in range(self.tp):
            shard = np.ascontiguousarray(w[:, r * sb:(r + 1) * sb]).reshape(-1)
            xs = np.ascontiguousarray(x[r * nc:(r + 1) * nc])
            if dt == orc.F16:
                part = self.o.gemv_f16(shard.view(np.float16).reshape(rows, nc), xs, rows, nc)
            elif dt in AX8:
                part = self.o.gemv_ax8(dt, shard, self.o.quantize(orc.Q8_B32T2, xs[None, :]), rows, nc)
            else:
                part = self.o.gemv_half(dt, shard, xs, rows, nc)
            part = orc.f16(part)
            acc = part if acc is None else orc.f16(acc + part)
        return orc.f16(acc + resid)

    def _rowpar_shard(self, name, dt, x, rows, cols, r):
        """rank r's partial product of a row-parallel matrix: round_h(W[:, shard r] x[shard r])"""
        rb = orc.row_bytes(dt, cols) if dt != orc.F16 else cols * 2
        w = self.t[name][: rb * rows].reshape(rows, rb)
        nc, sb = cols // self.tp, rb // self.tp
        shard = np.ascontiguousarray(w[:, r * sb:(r + 1) * sb]).reshape(-1)
        xs = np.ascontiguousarray(x[r * nc:(r + 1) * nc])
        if dt == orc.F16:
            return orc.f16(self.o.gemv_f16(shard.view(np.float16).reshape(rows, nc), xs, rows, nc))
        if dt in AX8:
            return orc.f16(self.o.gemv_ax8(dt, shard, self.o.quantize(orc.Q8_B32T2, xs[None, :]), rows, nc))
        return orc.f16(self.o.gemv_half(dt, shard, xs, rows, nc))

    def _ffn(self, prefix, wdt, xn, d, f):
        """GLU FFN on an already normalised input: silu(w1 x) * (w3 x) -> w2, every tensor stored as half"""
        g = orc.f16(self._gemv(prefix + "w1", wdt, xn, f, d))
        u = orc.f16(self._gemv(prefix + "w3", wdt, xn, f, d))
        h = orc.f16(orc.f16(g / (1.0 + np.exp(-g.astype(np.float64)))) * u)
        return orc.f16(self._gemv(prefix + "w2", wdt, h, d, f))

    @staticmethod
    def router_softmax(logits_f16):
        """Tensor_SoftMax_Alg2_Kernel (unary_tensor_opr.h:481) on one row of <= 32 router logits: fp32 max / exp /
        butterfly sum over a warp, exponentials stored as half, then multiplied by half(1/sum) in half"""
        v = np.zeros(32, dtype=np.float32)
        n = len(logits_f16)
        v[:n] = logits_f16.astype(np.float32)
        mx = np.float32(v[:n].max())
        ex = np.zeros(32, dtype=np.float32)
        ex[:n] = np.exp((v[:n] - mx).astype(np.float32)).astype(np.float32)
        ssum = ex.copy()
        idx = np.arange(32)
        for o in (16, 8, 4, 2, 1):
            ssum = (ssum + ssum[idx ^ o]).astype(np.float32)
        scale = np.float16(np.float32(1.0) / ssum[0])
        return (ex[:n].astype(np.float16).astype(np.float32) * np.float32(scale)).astype(np.float16)

    def _moe(self, p, wdt, x1n, x1, d, f):
        """ProcessGpuLayer_Moe (inference_worker.cc:1924-2146) for one token"""
        s, o = self.s, self.o
        gate = self.t[p + "gate"][: s.experts * d * 2].view(np.float16).reshape(s.experts, d)
        logits = orc.f16(o.gemv_f16(gate, x1n, s.experts, d)).astype(np.float16)
        probs = self.router_softmax(logits)
        ids, wts, cnt = o.moe_topk(probs.astype(np.float32)[None, :], s.experts_top_k, bool(s.moe_norm_top_k_prob))
        sel = sorted((int(ids[0, j]), float(wts[0, j])) for j in range(int(cnt[0])))     # ascending expert id
        self.last_route = sel
        if self.tp > 1:
            # tensor parallel (ours; the reference cannot): the experts' w2 are row-parallel, rank r folds ITS partial
            # outputs with the same half FMA chain, the ranks' vectors are added in rank order, the residual last
            acc = None
            for r in range(self.tp):
                out = np.zeros(d, dtype=np.float16)
                for e, w in sel:
                    pre = f"{p}experts.{e}."
                    g = orc.f16(self._gemv(pre + "w1", wdt, x1n, f, d))
                    u = orc.f16(self._gemv(pre + "w3", wdt, x1n, f, d))
                    h = orc.f16(orc.f16(g / (1.0 + np.exp(-g.astype(np.float64)))) * u)
                    y = self._rowpar_shard(pre + "w2", wdt, h, d, f, r)
                    out = (y.astype(np.float64) * np.float64(np.float16(w)) + out.astype(np.float64)).astype(np.float16)
                part = out.astype(np.float32)
                acc = part if acc is None else orc.f16(acc + part)
            return orc.f16(acc + x1)
        out = np.zeros(d, dtype=np.float16)
        for e, w in sel:
            y = self._ffn(f"{p}experts.{e}.", wdt, x1n, d, f)
            # AddByRowIdx_Kernel: B = __hfma(A, w, B) -- one rounding
            out = (y.astype(np.float64) * np.float64(np.float16(w)) + out.astype(np.float64)).astype(np.float16)
        return orc.f16(x1 + out.astype(np.float32))

    def _norm(self, name, x):
        w = self.t[name][: x.size * 2].view(np.float16).astype(np.float32)
        return orc.f16(self.o.rmsnorm(x, w, self.s.norm_eps))

    def step(self, token, pos):
        s, o, hd = self.s, self.o, self.hd
        d, wdt = s.d_model, s.weight_dtype
        emb = self.t["tok_embeddings"].view(np.float16).reshape(s.vocab, d)
        x = emb[token].astype(np.float32)
        qd, kvd = s.heads * hd, s.kv_heads * hd
        for l in range(s.layers):
            p = f"layers.{l}."
            xn = self._norm(p + "attention_norm", x)
            q = orc.f16(self._gemv(p + "wq", wdt, xn, qd, d))
            k = orc.f16(self._gemv(p + "wk", wdt, xn, kvd, d))
            v = orc.f16(self._gemv(p + "wv", wdt, xn, kvd, d))
            q = orc.f16(o.rope(q, s.heads, hd, pos, s.rope_theta, s.rope_order)).reshape(-1)
            k = orc.f16(o.rope(k, s.kv_heads, hd, pos, s.rope_theta, s.rope_order)).reshape(-1)
            if s.kv_dtype == orc.Q8_B32T2:
                k = o.dequantize(orc.Q8_B32T2, o.quantize(orc.Q8_B32T2, k[None, :]), kvd)[0]
                v = o.dequantize(orc.Q8_B32T2, o.quantize(orc.Q8_B32T2, v[None, :]), kvd)[0]
            assert len(self.K[l]) == pos
            self.K[l].append(k); self.V[l].append(v)
            a = orc.f16(o.attn_decode(q, np.stack(self.K[l]), np.stack(self.V[l]), s.heads, s.kv_heads, hd, 1.0, True))
            x1 = self._gemv_rowpar(p + "wo", wdt, a, d, qd, x)
            x1n = self._norm(p + "ffn_norm", x1)
            if getattr(s, "experts", 0):
                x = self._moe(p, wdt, x1n, x1, d, s.d_ffn)
                continue
            g = orc.f16(self._gemv(p + "w1", wdt, x1n, s.d_ffn, d))
            u = orc.f16(self._gemv(p + "w3", wdt, x1n, s.d_ffn, d))
            h = orc.f16(orc.f16(g / (1.0 + np.exp(-g.astype(np.float64)))) * u)
            x = self._gemv_rowpar(p + "w2", wdt, h, d, s.d_ffn, x1)
        xn = self._norm("output_norm", x)
        return self._gemv("output", s.lm_head_dtype, xn, s.vocab, d)
