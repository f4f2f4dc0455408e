// Per-GPU decode runtime for pre-norm llama-family decoders.
//
// Replaces GpuInferenceWorker::{Run, ProcessPreLayer, ProcessGpuLayer, ProcessGpuLayer_Attention,
// Attention_CalculateCurQKV, ProcessGpuLayer_FeedForward, ProcessPostLayer}
// (reference src/transformer/inference_worker.cc:234-977,983-1405,1407-1620,1726-1900,552-640)
// for batch-1 decode, and the tensor partitioner's placement rules
// (src/transformer/network_builder.cc:1373-1406: wq/wk/wv/w1/w3 BY_COL = split output rows,
// wo/w2 BY_ROW = split the contraction dim on block boundaries, norms DUP).
//
// Where the reference runs ~60 synchronous launches per layer on the default stream and
// creates/joins a thread per token, this runtime issues 6 fused kernels per layer, chained
// with programmatic dependent launch, and replays the whole token as ONE CUDA graph whose
// position / token inputs live in device memory.
#include <cstdio>
#include <string>
#include <vector>
#include <cstring>
#include <cstdlib>

#include "ifb_formats.cuh"
#include "ifb_kernels.h"
#include "ifb_step.h"
#include "../../include/inferflow_b200.h"

using namespace ifb;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct LayerW {
    DevBuf attn_norm, wq, wk, wv, wo, ffn_norm, w1, w2, w3;
    DevBuf kcache, vcache;
    // mixture of experts (spec.experts > 0): router + one GLU FFN per expert; tbl_* = device arrays of the experts'
    // weight pointers, read by the GEMV kernel with the expert id the routing kernel chose
    DevBuf gate;
    std::vector<DevBuf> ew1, ew3, ew2;
    DevBuf tbl_w1, tbl_w3, tbl_w2;
};

struct ifb_model {
    ifb_model_spec_t spec;
    int hd = 0;            // head_dim
    int heads_l = 0, kv_heads_l = 0, ffn_l = 0;   // this rank's shard
    int q_dim_l = 0, kv_dim_l = 0;
    std::vector<LayerW> layers;
    DevBuf tok_emb, out_norm, lm_head;
    DevBuf router, moe_ids, moe_wts, moe_y;   // MoE: router logits [experts] f16, slots' expert ids / weights, expert outputs [top_k][d]
    DevBuf xin;            // pipeline stage with layer_begin > 0: the hidden state handed over by the previous stage
    DevBuf x0, x1, qkv, attn_out, hbuf, logits, attn_ws, scalars, trace;   // trace: [512 launches][160 CTAs][8] ns stamps
    int trace_on = 0;
    DevBuf counters;   // one work cursor per launch of a step, zeroed at the start of every step   // scalars: [0]=pos, [1]=token, [2]=next
    std::vector<void *> allocs;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    bool graph_ready = false;
    int launches_per_step = 0;
    int use_pdl = 1;
    int use_graph = 1;
    int64_t weight_bytes = 0;
    // tensor-parallel exchange (peer memory): same layout on every rank
    DevBuf xchg;                    // this rank's exchange region
    std::vector<void *> peer_xchg;  // every rank's region as mapped here, [tp_size] (own entry = xchg.p)
    int64_t x_recv = 0, x_recv_parity = 0, x_recv_rank = 0;   // partial sums: [2 parities][8 ranks][d_model/2] words {half2, flag}
    int64_t x_logits = 0;                                     // logits of all ranks: [vocab/2] words {half2, flag}
    int vocab_l = 0;                // lm_head rows held by this rank (vocabulary-sharded under TP)
    // persistent step kernel (ifb_step.cu): ONE launch per token for dense models it supports
    int use_step = 1;               // IFB_STEP=0 keeps the multi-launch path
    bool step_built = false, step_ok = false;
    StepParams step;
    DevBuf step_layers, step_ctr, partials, ybuf, qkvw, attw, hbufw;
    int *step_err = nullptr;        // pinned host memory: error record of a step kernel that hit its deadline
    int host_pos = -1;              // host mirror of the device-side position (decode_next must not run past max_ctx)
    // prompt pass (ifb_model_prefill): activations of T tokens, grown on demand
    int pf_cap = 0;
    void *pf_x = nullptr, *pf_xn = nullptr, *pf_q = nullptr, *pf_k = nullptr, *pf_v = nullptr, *pf_att = nullptr, *pf_h1 = nullptr,
         *pf_h3 = nullptr, *pf_tmp = nullptr, *pf_tok = nullptr;
    // batched decode (ifb_model_decode_batch): several sequences, each with its own KV slot, one token each per step
    int kv_slots = 1;
    int64_t kv_slot_bytes = 0;      // bytes of one slot of a layer's K (or V) cache
    DevBuf bt_pos, bt_next, bt_logits;
    int api_wdt = 0, api_ldt = 0;   // element types the caller named (spec.weight_dtype / lm_head_dtype hold the STORAGE types)
    int l0 = 0, nl = 0;             // pipeline stage: this instance holds the global layers [l0, l0 + nl)
    bool first_stage = true, last_stage = true;
    int64_t smem_kb[5] = {0, 0, 0, 0, 0};   // per-CTA shared-memory budget of the QKV / wo / GLU / w2 / lm_head GEMVs (KB, 0 = kernel default)
};

static int dev_alloc(ifb_model *m, DevBuf &b, size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return IFB_ERR_CUDA;
    }
    b.p = p;
    b.bytes = bytes;
    m->allocs.push_back(p);
    return IFB_OK;
}

#define TRY(x)                      \
    do {                            \
        int rc_ = (x);              \
        if (rc_ != IFB_OK) return rc_; \
    } while (0)

static int64_t wbytes(int dt, int64_t rows, int64_t cols) { return rows * row_bytes(dt, cols); }

extern "C" int32_t ifb_model_create(ifb_model_t **out, const ifb_model_spec_t *spec) {
    IFB_CHECK_ARG(out != nullptr && spec != nullptr, "model_create: null argument");
    const ifb_model_spec_t &s = *spec;
    IFB_CHECK_ARG(s.layers > 0 && s.d_model > 0 && s.heads > 0 && s.kv_heads > 0 && s.vocab > 0,
                  "model_create: bad hyper-parameters");
    IFB_CHECK_ARG(s.d_model % s.heads == 0 && s.heads % s.kv_heads == 0, "heads do not divide d_model");
    const int tp = s.tp_size > 0 ? s.tp_size : 1;
    // network_builder.cc:1207: heads % ndev == 0 ; kv heads are sharded with them
    IFB_CHECK_ARG(s.heads % tp == 0 && s.kv_heads % tp == 0 && s.d_ffn % tp == 0,
                  "tensor parallelism %d does not divide heads %d / kv_heads %d / d_ffn %d", tp,
                  s.heads, s.kv_heads, s.d_ffn);
    IFB_CHECK_ARG(s.experts >= 0 && s.experts <= 32, "experts %d (the routing kernel handles up to 32)", s.experts);
    IFB_CHECK_ARG(s.experts == 0 || (s.experts_top_k >= 1 && s.experts_top_k <= 8 && s.experts_top_k <= s.experts),
                  "experts_top_k %d (RowItemForMoe::MAX_SIZE is 8)", s.experts_top_k);
    // MoE under tensor parallelism (the reference cannot: its expert tensors reach one device only): every
    // expert is split like the dense FFN -- w1 / w3 BY_COL, w2 BY_ROW -- the gate is duplicated
    IFB_CHECK_ARG(tp <= 8 && s.vocab % tp == 0 && (s.vocab / tp) % 8 == 0, "vocab %d cannot be sharded %d ways", s.vocab, tp);
    IFB_CHECK_ARG(tp == 1 || s.layers <= 127, "tensor parallel: at most 127 layers (exchange index is 8 bits)");
    // weight_repack: 3.5-bit blocks are stored as nibble blocks with the same codes / base / scale (ifb_quantize.cu)
    const int wdt = (s.weight_repack && s.weight_dtype == DT_Q3H_B64T1) ? DT_Q4_B64T1 : s.weight_dtype;
    const int ldt_store = (s.weight_repack && s.lm_head_dtype == DT_Q3H_B64T1) ? DT_Q4_B64T1 : s.lm_head_dtype;
    const int wcap = block_capacity(wdt);
    IFB_CHECK_ARG(wcap > 0 && wdt != DT_F32, "unsupported device_weight_data_type %d", wdt);
    // pipeline stage: layers [layer_begin, layer_end) of the model (0, 0 = all)
    const int lb = s.layer_begin, le = (s.layer_begin == 0 && s.layer_end == 0) ? s.layers : s.layer_end;
    IFB_CHECK_ARG(lb >= 0 && le > lb && le <= s.layers, "pipeline stage: layers [%d, %d) of %d", lb, le, s.layers);
    IFB_CHECK_ARG(s.kv_slots >= 0 && s.kv_slots <= 16, "kv_slots %d (at most 16 concurrent sequences, query_state_table.h)", s.kv_slots);
    IFB_CHECK_ARG((lb == 0 && le == s.layers) || tp == 1, "pipeline stages of a tensor-parallel group are not supported yet (stage [%d, %d), tp %d)", lb, le, tp);
    // BY_ROW split must fall on block boundaries (device_tensor_builder.cu:161-330)
    IFB_CHECK_ARG((s.d_model / tp) % (wcap > 32 ? wcap : 32) == 0 && (s.d_ffn / tp) % (wcap > 32 ? wcap : 32) == 0,
                  "BY_ROW shard (d_model/%d=%d, d_ffn/%d=%d) is not a multiple of the block capacity %d",
                  tp, s.d_model / tp, tp, s.d_ffn / tp, wcap);

    ifb_model *m = new ifb_model();
    m->spec = s;
    m->spec.tp_size = tp;
    m->api_wdt = s.weight_dtype; m->api_ldt = s.lm_head_dtype;
    m->spec.weight_dtype = wdt; m->spec.lm_head_dtype = ldt_store;       // from here on: the storage types
    m->kv_slots = s.kv_slots > 1 ? s.kv_slots : 1;
    m->l0 = lb; m->nl = le - lb;
    m->first_stage = lb == 0; m->last_stage = le == s.layers;
    m->hd = s.d_model / s.heads;
    m->heads_l = s.heads / tp;
    m->kv_heads_l = s.kv_heads / tp;
    m->ffn_l = s.d_ffn / tp;
    m->q_dim_l = m->heads_l * m->hd;
    m->kv_dim_l = m->kv_heads_l * m->hd;
    m->vocab_l = s.vocab / tp;      // ours: lm_head rows sharded (the reference keeps it whole on the last GPU,
                                    // network_builder.cc:1271-1283); the slices meet in every rank's exchange region
    const char *e1 = getenv("IFB_NO_PDL");
    if (e1 && atoi(e1)) m->use_pdl = 0;
    {
        // ring depth per launch kind (tuning aid; the defaults are set below from measurements on B200)
        const char *names[5] = {"IFB_SMEM_QKV", "IFB_SMEM_WO", "IFB_SMEM_GLU", "IFB_SMEM_W2", "IFB_SMEM_HEAD"};
        const int64_t defaults[5] = {150, 0, 0, 0, 0};   // QKV: leave room for the attention kernel's CTAs (early KV prefetch)
        for (int i = 0; i < 5; i++) {
            const char *e = getenv(names[i]);
            m->smem_kb[i] = e ? atoll(e) : defaults[i];
        }
    }
    const char *e2 = getenv("IFB_NO_GRAPH");
    if (e2 && atoi(e2)) m->use_graph = 0;
    // Which decode path: the persistent step kernel (one launch per token) or the multi-launch CUDA graph. Measured on
    // B200: on one and two GPUs the graph is 5-8 % ahead (both designs pay ~4 us per grid-wide
    // dependency: the step kernel in its op boundaries, the graph in launch + staging); from four GPUs on the step kernel
    // wins (per-rank streaming shrinks, launch overheads do not). IFB_STEP=0/1 overrides.
    m->use_step = (tp >= 4 && m->first_stage && m->last_stage) ? 1 : 0;       // measured crossover (7B Q4, ctx 2k): 1 GPU 556 vs 600, 2 GPUs 641 vs 675, 4 GPUs 716 vs 671 tok/s
    const char *e4 = getenv("IFB_STEP");
    if (e4) m->use_step = atoi(e4) ? 1 : 0;
    if (!(m->first_stage && m->last_stage)) m->use_step = 0;     // a pipeline stage runs the multi-launch path

    int rc = IFB_OK;
    auto A = [&](DevBuf &b, size_t bytes) { if (rc == IFB_OK) rc = dev_alloc(m, b, bytes); };
    const int d = s.d_model;
    m->layers.resize(m->nl);
    int64_t wsum = 0;
    for (auto &L : m->layers) {
        A(L.attn_norm, d * 2);
        A(L.ffn_norm, d * 2);
        A(L.wq, wbytes(wdt, m->q_dim_l, d));
        A(L.wk, wbytes(wdt, m->kv_dim_l, d));
        A(L.wv, wbytes(wdt, m->kv_dim_l, d));
        A(L.wo, wbytes(wdt, d, m->q_dim_l));
        const int64_t ffn_bytes = 2 * wbytes(wdt, m->ffn_l, d) + wbytes(wdt, d, m->ffn_l);
        if (s.experts == 0) {
            A(L.w1, wbytes(wdt, m->ffn_l, d));
            A(L.w3, wbytes(wdt, m->ffn_l, d));
            A(L.w2, wbytes(wdt, d, m->ffn_l));
            wsum += ffn_bytes;
        } else {
            A(L.gate, (size_t)s.experts * d * 2);     // < tensor_quant_threshold: stays fp16 (network_builder.cc:1558)
            L.ew1.resize(s.experts); L.ew3.resize(s.experts); L.ew2.resize(s.experts);
            std::vector<void *> p1(s.experts), p3(s.experts), p2(s.experts);
            for (int e = 0; e < s.experts; e++) {
                A(L.ew1[e], wbytes(wdt, m->ffn_l, d));
                A(L.ew3[e], wbytes(wdt, m->ffn_l, d));
                A(L.ew2[e], wbytes(wdt, d, m->ffn_l));
                p1[e] = L.ew1[e].p; p3[e] = L.ew3[e].p; p2[e] = L.ew2[e].p;
            }
            A(L.tbl_w1, s.experts * sizeof(void *));
            A(L.tbl_w3, s.experts * sizeof(void *));
            A(L.tbl_w2, s.experts * sizeof(void *));
            if (rc == IFB_OK) {
                cudaMemcpy(L.tbl_w1.p, p1.data(), s.experts * sizeof(void *), cudaMemcpyHostToDevice);
                cudaMemcpy(L.tbl_w3.p, p3.data(), s.experts * sizeof(void *), cudaMemcpyHostToDevice);
                cudaMemcpy(L.tbl_w2.p, p2.data(), s.experts * sizeof(void *), cudaMemcpyHostToDevice);
            }
            wsum += (int64_t)s.experts * d * 2 + (int64_t)s.experts_top_k * ffn_bytes;   // bytes READ per token
        }
        wsum += wbytes(wdt, m->q_dim_l, d) + 2 * wbytes(wdt, m->kv_dim_l, d) + wbytes(wdt, d, m->q_dim_l);
        int64_t kvb = kv_cache_bytes(s.kv_dtype, m->kv_heads_l, m->hd, s.max_ctx);
        if (kvb < 0 && rc == IFB_OK) { set_last_error("unsupported kv cache type %d", s.kv_dtype); rc = IFB_ERR_TYPE; }
        kvb = (kvb + 255) / 256 * 256;
        m->kv_slot_bytes = kvb;
        A(L.kcache, kvb * m->kv_slots);      // slot 0 is the cache of the single-sequence paths
        A(L.vcache, kvb * m->kv_slots);
    }
    A(m->tok_emb, m->first_stage ? (size_t)s.vocab * d * 2 : 256);      // the embedding lives on the first stage,
    A(m->out_norm, d * 2);
    const int ldt = ldt_store;
    if (rc == IFB_OK && (block_capacity(ldt) <= 0 || ldt == DT_F32)) {
        set_last_error("unsupported lm_head type %d", ldt);
        rc = IFB_ERR_TYPE;
    }
    A(m->lm_head, m->last_stage ? wbytes(ldt, m->vocab_l, d) : 256);     // the lm_head on the last
    if (m->last_stage) wsum += wbytes(ldt, m->vocab_l, d);
    {   // the exchange region: under tensor parallelism the peers store into it; on one GPU the step kernel's own
        // row-parallel GEMVs do (flag-in-data words instead of a completion count, see ifb_step.cu)
        auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
        m->x_recv = 0;
        m->x_recv_rank = (int64_t)d * 4;             // d/2 words of 8 bytes
        m->x_recv_parity = 8 * m->x_recv_rank;
        m->x_logits = up(2 * m->x_recv_parity);
        A(m->xchg, up(m->x_logits + (int64_t)s.vocab * 4));
        if (rc == IFB_OK && cudaMemset(m->xchg.p, 0, m->xchg.bytes) != cudaSuccess) { set_last_error("cudaMemset failed"); rc = IFB_ERR_CUDA; }
    }
    m->weight_bytes = wsum;
    A(m->x0, d * 2);
    A(m->x1, d * 2);
    A(m->xin, d * 2);
    A(m->qkv, (size_t)(m->q_dim_l + 2 * m->kv_dim_l) * 2);
    A(m->attn_out, (size_t)m->q_dim_l * 2);
    A(m->hbuf, (size_t)m->ffn_l * 2);
    if (s.experts > 0) {
        A(m->router, 256);
        A(m->moe_ids, 64);
        A(m->moe_wts, 64);
        A(m->moe_y, (size_t)s.experts_top_k * d * 2);
    }
    A(m->logits, (size_t)s.vocab * 2);
    if (m->kv_slots > 1) {
        A(m->bt_pos, 16 * 4);
        A(m->bt_next, 16 * 4);
        A(m->bt_logits, (size_t)m->kv_slots * s.vocab * 2);
    }
    A(m->attn_ws, attn_workspace_bytes(m->heads_l, m->hd));
    A(m->scalars, 256);
    {   // one 8-int cursor block per launch of a step: dense 5 per layer, MoE 6 + 2*top_k per layer, + embedding, lm_head, argmax, advance
        const size_t per_layer = s.experts > 0 ? 6 + 2 * (size_t)s.experts_top_k : 5;
        A(m->counters, (per_layer * m->nl + 8) * 8 * sizeof(int));
    }
    A(m->ybuf, d * 2);
    A(m->qkvw, (size_t)(m->q_dim_l + 2 * m->kv_dim_l) * 4);     // {half2, flag} words
    A(m->attw, (size_t)m->q_dim_l * 4);
    A(m->hbufw, (size_t)m->ffn_l * 4);
    if (rc == IFB_OK) { cudaMemset(m->qkvw.p, 0, m->qkvw.bytes); cudaMemset(m->attw.p, 0, m->attw.bytes); cudaMemset(m->hbufw.p, 0, m->hbufw.bytes); }
    A(m->partials, (size_t)m->heads_l * ((s.max_ctx + 31) / 32) * (m->hd + 4) * sizeof(float));
    { const char *e3 = getenv("IFB_TRACE"); if (e3 && atoi(e3)) m->trace_on = 1; }
    if (m->trace_on) A(m->trace, (size_t)512 * 160 * 8 * 8);
    if (rc == IFB_OK && m->trace_on) cudaMemset(m->trace.p, 0, m->trace.bytes);
    if (rc == IFB_OK) {
        if (cudaMemset(m->attn_ws.p, 0, m->attn_ws.bytes) != cudaSuccess ||
            cudaMemset(m->scalars.p, 0, 256) != cudaSuccess) {
            set_last_error("cudaMemset failed");
            rc = IFB_ERR_CUDA;
        }
        for (auto &L : m->layers) {
            cudaMemset(L.kcache.p, 0, L.kcache.bytes);
            cudaMemset(L.vcache.p, 0, L.vcache.bytes);
        }
    }
    if (rc != IFB_OK) {
        ifb_model_destroy(m);
        return rc;
    }
    *out = m;
    return IFB_OK;
}

extern "C" void ifb_model_destroy(ifb_model_t *m) {
    if (!m) return;
    if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
    if (m->graph) cudaGraphDestroy(m->graph);
    for (size_t p = 0; p < m->peer_xchg.size(); p++)       // unmap the peers' exchange regions (ours is in allocs)
        if ((int)p != m->spec.tp_rank && m->peer_xchg[p] != nullptr) cudaIpcCloseMemHandle(m->peer_xchg[p]);
    for (void *p : m->allocs) cudaFree(p);
    if (m->step_err) cudaFreeHost(m->step_err);
    for (void *p : {m->pf_x, m->pf_xn, m->pf_q, m->pf_k, m->pf_v, m->pf_att, m->pf_h1, m->pf_h3, m->pf_tmp, m->pf_tok})
        if (p) cudaFree(p);
    delete m;
}

extern "C" int64_t ifb_model_weight_bytes(const ifb_model_t *m) { return m ? m->weight_bytes : 0; }
extern "C" int32_t ifb_model_storage_dtype(const ifb_model_t *m, int32_t which) {
    if (!m) return -1;
    return which == 1 ? m->spec.lm_head_dtype : m->spec.weight_dtype;
}
extern "C" void *ifb_model_hidden_ptr(ifb_model_t *m, int32_t which) {
    if (!m) return nullptr;
    return which == 0 ? m->xin.p : m->x0.p;      // 0: where a later stage expects its input; 1: where a stage leaves its output
}

extern "C" int64_t ifb_model_kv_bytes_per_token(const ifb_model_t *m, int32_t ctx_len) {
    if (!m) return 0;
    // 2 * L * ctx * (g*e) * kvb   (SURVEY section 8d), this rank's shard
    double kvb = m->spec.kv_dtype == DT_F16 ? 2.0 : 34.0 / 32.0;
    return (int64_t)(2.0 * m->nl * (double)ctx_len * m->kv_dim_l * kvb);
}

// ---------------------------------------------------------------------------------------
// synthetic weights: generated on the device directly in block form
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float gauss(uint64_t key) {
    float u1 = u01(mix64(key)), u2 = u01(mix64(key ^ 0xabcdef1234567ull));
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// levels-1 of each format (largest code)
__device__ __forceinline__ int max_code(int dt) {
    switch (dt) {
        case DT_Q8_B32T1: return 255;
        case DT_Q6_B64T1: return 63;
        case DT_Q5_B64T1: case DT_Q5_B32T1: return 31;
        case DT_Q4_B16: case DT_Q4_B32T1A: case DT_Q4_B32T1B: case DT_Q4_B64T1: return 15;
        case DT_Q3H_B64T1: return 10;
        case DT_Q3_B32T1A: case DT_Q3_B32T1B: return 7;
        case DT_Q2_B32T1A: case DT_Q2_B32T1B: return 3;
        default: return 127;
    }
}

// Fill a block-quantized tensor: random code bytes, scale ~ |N(0, sigma)|*2/maxcode, base = -maxcode/2*scale
// (so dequantized weights are ~uniform in +-|N(0,sigma)|, sigma = 0.02 like N(0,0.02) weights).
// Q3H needs valid base-11 pairs (P <= 120): its code bytes are generated through the quantizer's packing.
// Every block is a function of (seed, GLOBAL block index of the unsharded tensor): a tensor-parallel rank that holds rows
// [row0, ...) and blocks [blk0, blk0 + lbpr) of every row generates exactly the bytes a single-GPU model holds there, so a
// sharded run can be checked against an unsharded one (bench.py does, before it times anything).
__global__ void synth_blocks_kernel(uint8_t *dst, int dt, int64_t nblocks, int bb, uint64_t seed, float sigma,
                                    int64_t lbpr, int64_t gbpr, int64_t row0, int64_t blk0) {
    int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nblocks) return;
    uint8_t *b = dst + li * bb;
    const int64_t bi = (row0 + li / lbpr) * gbpr + blk0 + li % lbpr;
    uint64_t key = seed * 0x100000001b3ull + (uint64_t)bi;
    float amp = fabsf(gauss(key)) * sigma + 1e-4f;
    int mc = max_code(dt);
    if (dt == DT_Q8_B32T2) {
        *reinterpret_cast<uint16_t *>(b) = f2h_bits(amp / 127.0f);   // one 16-bit store (see ifb_quantize.cu put16)
        for (int i = 0; i < 32; i += 8) {
            uint64_t r = mix64(key * 31 + i);
            for (int k = 0; k < 8; k++) b[2 + i + k] = (uint8_t)(r >> (8 * k));
        }
        return;
    }
    float scale = 2.0f * amp / (float)mc, base = -amp;
    int hdr = 4;
    if (dt == DT_Q4_B16) {
        b[0] = (uint8_t)(base * 100 + 100.5f); b[1] = (uint8_t)(scale * 1000 + 0.5f); hdr = 2;
    } else if (dt == DT_Q5_B32T1) {
        *reinterpret_cast<uint16_t *>(b) = f2h_bits(scale);
        *reinterpret_cast<uint16_t *>(b + 2) = f2h_bits(base);
    } else {
        *reinterpret_cast<uint16_t *>(b) = f2h_bits(base);
        *reinterpret_cast<uint16_t *>(b + 2) = f2h_bits(scale);
    }
    if (dt == DT_Q3H_B64T1) {
        for (int i = 4; i < 8; i++) b[i] = 0;
        for (int r = 0; r < 8; r++) {
            uint64_t rr = mix64(key * 131 + r);
            int P[4];
            for (int k = 0; k < 4; k++) P[k] = (int)((rr >> (16 * k)) & 0xffff) % 121;
            b[16 + 2 * r] = (uint8_t)((P[0] & 15) | ((P[1] & 15) << 4));
            b[16 + 2 * r + 1] = (uint8_t)((P[2] & 15) | ((P[3] & 15) << 4));
            b[8 + r] = (uint8_t)(((P[0] >> 4) & 3) | (((P[1] >> 4) & 3) << 2) | (((P[2] >> 4) & 3) << 4) | (((P[3] >> 4) & 3) << 6));
            int hn = ((P[0] >> 6) & 1) | (((P[1] >> 6) & 1) << 1) | (((P[2] >> 6) & 1) << 2) | (((P[3] >> 6) & 1) << 3);
            b[4 + r / 2] |= (uint8_t)(hn << ((r & 1) * 4));
        }
        return;
    }
    for (int i = hdr; i < bb; i += 8) {
        uint64_t r = mix64(key * 31 + i);
        for (int k = 0; k < 8 && i + k < bb; k++) b[i + k] = (uint8_t)(r >> (8 * k));
    }
}

__global__ void synth_f16_kernel(__half *dst, int64_t n, uint64_t seed, float sigma, float constant, int use_const,
                                 int64_t lcols, int64_t gcols, int64_t row0, int64_t col0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t gi = (row0 + i / lcols) * gcols + col0 + i % lcols;
    dst[i] = __float2half_rn(use_const ? constant : gauss(seed * 0x100000001b3ull + (uint64_t)gi) * sigma);
}

// this rank's [rows x cols] shard at (row0, col0) of a global tensor with gcols columns
// blocks made in the caller's element type `api_dt` by make(buffer), kept in the storage type `store_dt`
// (the same unless weight_repack turned 3.5-bit blocks into nibble blocks)
template <class F>
static int produce_blocks(void *dst, int store_dt, int api_dt, int64_t rows, int64_t cols, cudaStream_t st, F make) {
    if (store_dt == api_dt) return make(dst);
    IFB_CHECK_ARG(api_dt == DT_Q3H_B64T1 && store_dt == DT_Q4_B64T1, "no repacking from type %d to type %d", api_dt, store_dt);
    void *tmp = nullptr;
    IFB_CUDA(cudaMalloc(&tmp, (size_t)(rows * row_bytes(api_dt, cols))));
    int rc = make(tmp);
    if (rc == IFB_OK) rc = repack_q3h_q4b64(dst, tmp, rows * (cols / 64), st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
}

static int synth_tensor(void *p_store, int store_dt, int dt, int64_t rows, int64_t cols, uint64_t seed, float sigma, cudaStream_t st,
                        int64_t gcols = 0, int64_t row0 = 0, int64_t col0 = 0) {
    if (gcols == 0) gcols = cols;
    return produce_blocks(p_store, store_dt, dt, rows, cols, st, [&](void *p) -> int {
    if (dt == DT_F16) {
        int64_t n = rows * cols;
        synth_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((__half *)p, n, seed, sigma, 0.f, 0, cols, gcols, row0, col0);
    } else {
        int cap = block_capacity(dt), bb = block_bytes(dt);
        int64_t nb = rows * (cols / cap);
        synth_blocks_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((uint8_t *)p, dt, nb, bb, seed, sigma, cols / cap, gcols / cap, row0, col0 / cap);
    }
    IFB_LAUNCH_CHECK();
    return IFB_OK;
    });
}
static int const_f16(void *p, int64_t n, float v, cudaStream_t st) {
    synth_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((__half *)p, n, 0, 0.f, v, 1, n, n, 0, 0);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

extern "C" int32_t ifb_model_init_synthetic(ifb_model_t *m, uint64_t seed, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr, "null model");
    cudaStream_t st = (cudaStream_t)stream;
    const auto &s = m->spec;
    const int d = s.d_model, wdt = s.weight_dtype;
    // every shard is generated from (global seed, tensor id, GLOBAL element position): the ranks of a tensor-parallel
    // group together hold exactly the single-GPU model of the same seed
    uint64_t sd = seed * 1000003ull;
    const int64_t rk = s.tp_rank;
    const int64_t q0 = rk * m->q_dim_l, kv0 = rk * m->kv_dim_l, f0 = rk * m->ffn_l;
    const int64_t gq = (int64_t)s.heads * m->hd, gf = s.d_ffn;
    int li = m->l0;                  // GLOBAL layer index: a pipeline stage holds its slice of the same model
    for (auto &L : m->layers) {
        uint64_t ls = sd + 100ull * (uint64_t)(li + 1);
        TRY(const_f16(L.attn_norm.p, d, 1.0f, st));
        TRY(const_f16(L.ffn_norm.p, d, 1.0f, st));
        TRY(synth_tensor(L.wq.p, wdt, m->api_wdt, m->q_dim_l, d, ls + 1, 0.02f, st, d, q0, 0));          // BY_COL: rows of this rank
        TRY(synth_tensor(L.wk.p, wdt, m->api_wdt, m->kv_dim_l, d, ls + 2, 0.02f, st, d, kv0, 0));
        TRY(synth_tensor(L.wv.p, wdt, m->api_wdt, m->kv_dim_l, d, ls + 3, 0.02f, st, d, kv0, 0));
        TRY(synth_tensor(L.wo.p, wdt, m->api_wdt, d, m->q_dim_l, ls + 4, 0.02f, st, gq, 0, q0));          // BY_ROW: columns of this rank
        if (s.experts == 0) {
            TRY(synth_tensor(L.w1.p, wdt, m->api_wdt, m->ffn_l, d, ls + 5, 0.02f, st, d, f0, 0));
            TRY(synth_tensor(L.w3.p, wdt, m->api_wdt, m->ffn_l, d, ls + 6, 0.02f, st, d, f0, 0));
            TRY(synth_tensor(L.w2.p, wdt, m->api_wdt, d, m->ffn_l, ls + 7, 0.02f, st, gf, 0, f0));
        } else {
            TRY(synth_tensor(L.gate.p, DT_F16, DT_F16, s.experts, d, ls + 8, 0.05f, st));   // router: non-degenerate expert choice
            for (int e = 0; e < s.experts; e++) {
                TRY(synth_tensor(L.ew1[e].p, wdt, m->api_wdt, m->ffn_l, d, ls + 20 + 3 * e, 0.02f, st, d, f0, 0));
                TRY(synth_tensor(L.ew3[e].p, wdt, m->api_wdt, m->ffn_l, d, ls + 21 + 3 * e, 0.02f, st, d, f0, 0));
                TRY(synth_tensor(L.ew2[e].p, wdt, m->api_wdt, d, m->ffn_l, ls + 22 + 3 * e, 0.02f, st, gf, 0, f0));
            }
        }
        li++;
    }
    if (m->first_stage) TRY(synth_tensor(m->tok_emb.p, DT_F16, DT_F16, s.vocab, d, seed * 1000003ull + 11, 1.0f, st));   // same on all ranks
    TRY(const_f16(m->out_norm.p, d, 1.0f, st));
    if (m->last_stage) TRY(synth_tensor(m->lm_head.p, s.lm_head_dtype, m->api_ldt, m->vocab_l, d, seed * 1000003ull + 13, 0.02f, st, d, rk * m->vocab_l, 0));   // vocabulary-sharded
    return IFB_OK;
}

// ---------------------------------------------------------------------------------------
// KV reset / synthetic fill
// ---------------------------------------------------------------------------------------
__global__ void synth_kv_kernel(uint8_t *cache, int kv_dtype, int kv_heads, int hd, int max_ctx_pad,
                                int ctx, uint64_t seed, int head0) {
    // one warp per (head, pos, block of 32 dims)
    int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int nb = hd / 32;
    int64_t total = (int64_t)kv_heads * ctx * nb;
    if (wid >= total) return;
    int blk = (int)(wid % nb);
    int pos = (int)((wid / nb) % ctx);
    int h = (int)(wid / ((int64_t)nb * ctx));
    // keyed by the GLOBAL kv head: the shards of a tensor-parallel group hold the single-GPU cache of the same seed
    const int64_t gwid = ((int64_t)(head0 + h) * ctx + pos) * nb + blk;
    float v = round_h(gauss(seed + (uint64_t)gwid * 32 + lane));
    int64_t base = (int64_t)h * max_ctx_pad + pos;
    if (kv_dtype == DT_F16) {
        reinterpret_cast<__half *>(cache)[base * hd + blk * 32 + lane] = __float2half_rn(v);
    } else {
        float mx = warp_max(fabsf(v));
        float scale = __fdiv_rn(mx, 127.0f);
        int q = scale <= 0.000001f ? 0 : (int)roundf(__fdiv_rn(v, scale));
        q = min(max(q, -128), 127);
        reinterpret_cast<int8_t *>(cache)[base * hd + blk * 32 + lane] = (int8_t)q;
        if (lane == 0)
            reinterpret_cast<__half *>(cache + (int64_t)kv_heads * max_ctx_pad * hd)[base * nb + blk] = __float2half_rn(scale);
    }
}

extern "C" int32_t ifb_model_kv_reset(ifb_model_t *m, int32_t ctx_len, int32_t fill, uint64_t seed,
                                      ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr, "null model");
    IFB_CHECK_ARG(ctx_len >= 0 && ctx_len <= m->spec.max_ctx, "kv_reset: ctx_len %d outside [0, %d]", ctx_len, m->spec.max_ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (fill && ctx_len > 0) {
        int pad = (m->spec.max_ctx + 63) / 64 * 64;
        int li = m->l0;              // GLOBAL layer index
        for (auto &L : m->layers) {
            int64_t warps = (int64_t)m->kv_heads_l * ctx_len * (m->hd / 32);
            unsigned grid = (unsigned)((warps * 32 + 255) / 256);
            const int head0 = m->spec.tp_rank * m->kv_heads_l;
            synth_kv_kernel<<<grid, 256, 0, st>>>((uint8_t *)L.kcache.p, m->spec.kv_dtype, m->kv_heads_l, m->hd, pad, ctx_len, seed + 977ull * li + 1, head0);
            synth_kv_kernel<<<grid, 256, 0, st>>>((uint8_t *)L.vcache.p, m->spec.kv_dtype, m->kv_heads_l, m->hd, pad, ctx_len, seed + 977ull * li + 500009, head0);
            li++;
        }
        IFB_LAUNCH_CHECK();
    }
    return IFB_OK;
}

// ---------------------------------------------------------------------------------------
// tensor loading (host data -> fp16 -> quantize on device -> keep this rank's shard)
// ---------------------------------------------------------------------------------------
__global__ void f32_to_f16_kernel(__half *dst, const float *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

// Upload rows [r0, r0+nr) x cols [c0, c0+nc) of a host [rows][cols] matrix as fp16, then quantize to dt.
static int load_shard(DevBuf &dst, int store_dt, int dt, const void *host, int host_dt, int64_t rows, int64_t cols,
                      int64_t r0, int64_t nr, int64_t c0, int64_t nc, cudaStream_t st) {
    const size_t esz = host_dt == DT_F32 ? 4 : 2;
    IFB_CHECK_ARG(host_dt == DT_F32 || host_dt == DT_F16, "load_tensor: host data must be F32 or F16");
    IFB_CHECK_ARG(wbytes(store_dt, nr, nc) <= (int64_t)dst.bytes, "load_tensor: shard %lldx%lld does not fit the device tensor", (long long)nr, (long long)nc);
    void *staging = nullptr, *half_buf = nullptr;
    IFB_CUDA(cudaMalloc(&staging, (size_t)nr * nc * esz));
    cudaError_t e = cudaMemcpy2DAsync(staging, nc * esz, (const uint8_t *)host + ((size_t)r0 * cols + c0) * esz,
                                      cols * esz, nc * esz, nr, cudaMemcpyHostToDevice, st);
    int rc = IFB_OK;
    if (e != cudaSuccess) { set_last_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = IFB_ERR_CUDA; }
    const void *h16 = staging;
    if (rc == IFB_OK && host_dt == DT_F32) {
        if (cudaMalloc(&half_buf, (size_t)nr * nc * 2) != cudaSuccess) { set_last_error("cudaMalloc failed"); rc = IFB_ERR_CUDA; }
        else {
            int64_t n = nr * nc;
            f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((__half *)half_buf, (const float *)staging, n);
            h16 = half_buf;
        }
    }
    if (rc == IFB_OK) {
        if (dt == DT_F16) {
            if (cudaMemcpyAsync(dst.p, h16, (size_t)nr * nc * 2, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { set_last_error("D2D copy failed"); rc = IFB_ERR_CUDA; }
        } else {
            rc = produce_blocks(dst.p, store_dt, dt, nr, nc, st, [&](void *p) -> int { return quantize_rows(dt, p, h16, nr, nc, st); });
        }
    }
    cudaStreamSynchronize(st);
    cudaFree(staging);
    if (half_buf) cudaFree(half_buf);
    return rc;
}

extern "C" int32_t ifb_model_load_tensor(ifb_model_t *m, const char *name, const void *host, int32_t host_dt,
                                         int64_t rows, int64_t cols, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr && name != nullptr && host != nullptr, "load_tensor: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const auto &s = m->spec;
    const int tp = s.tp_size, rk = s.tp_rank, d = s.d_model, wdt = s.weight_dtype;
    std::string n(name);
    // weights are read through fixed pointers: a captured graph stays valid
    if (n == "tok_embeddings" && !m->first_stage) return IFB_OK;                       // held by the first pipeline stage
    if ((n == "output_norm" || n == "output") && !m->last_stage) return IFB_OK;       // held by the last
    if (n == "tok_embeddings") return load_shard(m->tok_emb, DT_F16, DT_F16, host, host_dt, rows, cols, 0, rows, 0, cols, st);
    if (n == "output_norm") return load_shard(m->out_norm, DT_F16, DT_F16, host, host_dt, rows, cols, 0, rows, 0, cols, st);
    if (n == "output") {
        IFB_CHECK_ARG(rows == s.vocab && cols == d, "output shape");
        return load_shard(m->lm_head, s.lm_head_dtype, m->api_ldt, host, host_dt, rows, cols, (int64_t)rk * m->vocab_l, m->vocab_l, 0, cols, st);
    }
    int li = -1;
    char tn[64] = {0};
    if (sscanf(name, "layers.%d.%63s", &li, tn) != 2 || li < 0 || li >= s.layers) {
        set_last_error("load_tensor: unknown tensor '%s'", name);
        return IFB_ERR_ARG;
    }
    if (li < m->l0 || li >= m->l0 + m->nl) return IFB_OK;     // another pipeline stage holds this layer
    LayerW &L = m->layers[li - m->l0];
    std::string t(tn);
    const int hd = m->hd;
    // DUP
    if (t == "attention_norm") return load_shard(L.attn_norm, DT_F16, DT_F16, host, host_dt, rows, cols, 0, rows, 0, cols, st);
    if (t == "ffn_norm") return load_shard(L.ffn_norm, DT_F16, DT_F16, host, host_dt, rows, cols, 0, rows, 0, cols, st);
    // BY_COL (split output rows; heads stay whole)
    if (t == "wq") { IFB_CHECK_ARG(rows == (int64_t)s.heads * hd && cols == d, "wq shape"); return load_shard(L.wq, wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->q_dim_l, m->q_dim_l, 0, cols, st); }
    if (t == "wk") { IFB_CHECK_ARG(rows == (int64_t)s.kv_heads * hd && cols == d, "wk shape"); return load_shard(L.wk, wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->kv_dim_l, m->kv_dim_l, 0, cols, st); }
    if (t == "wv") { IFB_CHECK_ARG(rows == (int64_t)s.kv_heads * hd && cols == d, "wv shape"); return load_shard(L.wv, wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->kv_dim_l, m->kv_dim_l, 0, cols, st); }
    if (t == "w1") { IFB_CHECK_ARG(rows == s.d_ffn && cols == d, "w1 shape"); return load_shard(L.w1, wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->ffn_l, m->ffn_l, 0, cols, st); }
    if (t == "w3") { IFB_CHECK_ARG(rows == s.d_ffn && cols == d, "w3 shape"); return load_shard(L.w3, wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->ffn_l, m->ffn_l, 0, cols, st); }
    // BY_ROW (split the contraction dimension)
    if (t == "wo") { IFB_CHECK_ARG(rows == d && cols == (int64_t)s.heads * hd, "wo shape"); return load_shard(L.wo, wdt, m->api_wdt, host, host_dt, rows, cols, 0, rows, (int64_t)rk * m->q_dim_l, m->q_dim_l, st); }
    if (t == "w2") { IFB_CHECK_ARG(rows == d && cols == s.d_ffn, "w2 shape"); return load_shard(L.w2, wdt, m->api_wdt, host, host_dt, rows, cols, 0, rows, (int64_t)rk * m->ffn_l, m->ffn_l, st); }
    // mixture of experts: "gate" [experts][d] (router, kept fp16), "experts.E.w1|w2|w3"
    if (s.experts > 0) {
        if (t == "gate") { IFB_CHECK_ARG(rows == s.experts && cols == d, "gate shape"); return load_shard(L.gate, DT_F16, DT_F16, host, host_dt, rows, cols, 0, rows, 0, cols, st); }
        int e = -1; char wn[16] = {0};
        if (sscanf(tn, "experts.%d.%15s", &e, wn) == 2 && e >= 0 && e < s.experts) {
            std::string w(wn);
            if (w == "w1") { IFB_CHECK_ARG(rows == s.d_ffn && cols == d, "expert w1 shape"); return load_shard(L.ew1[e], wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->ffn_l, m->ffn_l, 0, cols, st); }
            if (w == "w3") { IFB_CHECK_ARG(rows == s.d_ffn && cols == d, "expert w3 shape"); return load_shard(L.ew3[e], wdt, m->api_wdt, host, host_dt, rows, cols, (int64_t)rk * m->ffn_l, m->ffn_l, 0, cols, st); }
            if (w == "w2") { IFB_CHECK_ARG(rows == d && cols == s.d_ffn, "expert w2 shape"); return load_shard(L.ew2[e], wdt, m->api_wdt, host, host_dt, rows, cols, 0, rows, (int64_t)rk * m->ffn_l, m->ffn_l, st); }
        }
    }
    set_last_error("load_tensor: unknown tensor '%s'", name);
    (void)tp;
    return IFB_ERR_ARG;
}

// ---------------------------------------------------------------------------------------
// llama2.c model files (the format of the reference's tiny models, model_reader.cc:3245-3427)
// ---------------------------------------------------------------------------------------
struct L2cHeader { int32_t dim, hidden, layers, heads, kv_heads, vocab, seq; bool shared; long data_off; };

static int l2c_read_header(FILE *f, L2cHeader &h) {
    uint32_t magic = 0;
    if (fread(&magic, 4, 1, f) != 1) { set_last_error("llama2.c: cannot read the file header"); return IFB_ERR_IO; }
    int version = 0;
    if (magic == 0x616b3432u) {   // "ak42"
        if (fread(&version, 4, 1, f) != 1 || version != 1) { set_last_error("llama2.c: version %d is not supported", version); return IFB_ERR_IO; }
    } else {
        fseek(f, 0, SEEK_SET);
    }
    int32_t v[7];
    if (fread(v, 4, 7, f) != 7) { set_last_error("llama2.c: truncated header"); return IFB_ERR_IO; }
    h.dim = v[0]; h.hidden = v[1]; h.layers = v[2]; h.heads = v[3]; h.kv_heads = v[4]; h.vocab = v[5]; h.seq = v[6];
    h.shared = h.vocab >= 0;
    if (h.vocab < 0) h.vocab = -h.vocab;
    if (version == 1) {           // v1: shared_classifier byte, then padding up to 256 bytes
        uint8_t sc = 0;
        if (fread(&sc, 1, 1, f) != 1) { set_last_error("llama2.c: truncated v1 header"); return IFB_ERR_IO; }
        h.shared = sc != 0;
        fseek(f, 256, SEEK_SET);
    }
    h.data_off = ftell(f);
    if (h.dim <= 0 || h.hidden <= 0 || h.layers <= 0 || h.heads <= 0 || h.kv_heads <= 0 || h.vocab <= 0 || h.seq <= 0 ||
        h.dim % h.heads != 0 || h.heads % h.kv_heads != 0 || h.dim > (1 << 20) || h.layers > 4096 || h.vocab > (1 << 24)) {
        set_last_error("llama2.c: implausible header (dim %d hidden %d layers %d heads %d kv_heads %d vocab %d seq %d)",
                       h.dim, h.hidden, h.layers, h.heads, h.kv_heads, h.vocab, h.seq);
        return IFB_ERR_IO;
    }
    return IFB_OK;
}

extern "C" int32_t ifb_llama2c_probe(const char *path, int32_t header_out[7], int32_t *shared_classifier) {
    IFB_CHECK_ARG(path != nullptr, "llama2c_probe: null path");
    FILE *f = fopen(path, "rb");
    if (!f) { set_last_error("llama2.c: cannot open %s", path); return IFB_ERR_IO; }
    L2cHeader h;
    int rc = l2c_read_header(f, h);
    if (rc == IFB_OK) {
        // the file must hold every tensor the header promises
        const int64_t hd = h.dim / h.heads, kvd = (int64_t)h.kv_heads * hd;
        int64_t floats = (int64_t)h.vocab * h.dim + (int64_t)h.layers * (2 * h.dim + 2 * (int64_t)h.dim * h.dim + 2 * kvd * h.dim +
                                                                         3 * (int64_t)h.hidden * h.dim) + h.dim;
        int64_t need = h.data_off + floats * 4 + (int64_t)h.seq * hd + (h.shared ? 0 : (int64_t)h.vocab * h.dim * 4);
        fseek(f, 0, SEEK_END);
        const int64_t have = ftell(f);
        if (have < need) { set_last_error("llama2.c: %s holds %lld bytes, the header needs %lld", path, (long long)have, (long long)need); rc = IFB_ERR_IO; }
    }
    fclose(f);
    if (rc != IFB_OK) return rc;
    if (header_out) { header_out[0] = h.dim; header_out[1] = h.hidden; header_out[2] = h.layers; header_out[3] = h.heads;
                      header_out[4] = h.kv_heads; header_out[5] = h.vocab; header_out[6] = h.seq; }
    if (shared_classifier) *shared_classifier = h.shared ? 1 : 0;
    return IFB_OK;
}

extern "C" int32_t ifb_model_load_llama2c(ifb_model_t **out, const char *path, int32_t weight_dtype, int32_t kv_dtype,
                                          int32_t max_ctx, int32_t rope_order, ifb_stream_t stream) {
    return ifb_model_load_llama2c_stage(out, path, weight_dtype, kv_dtype, max_ctx, rope_order, 0, 0, stream);
}

extern "C" int32_t ifb_model_load_llama2c_stage(ifb_model_t **out, const char *path, int32_t weight_dtype, int32_t kv_dtype,
                                                int32_t max_ctx, int32_t rope_order, int32_t layer_begin, int32_t layer_end,
                                                ifb_stream_t stream) {
    IFB_CHECK_ARG(out != nullptr && path != nullptr, "load_llama2c: null argument");
    int32_t hv[7];
    int32_t shared = 1;
    TRY(ifb_llama2c_probe(path, hv, &shared));
    ifb_model_spec_t s = {};
    s.d_model = hv[0]; s.d_ffn = hv[1]; s.layers = hv[2]; s.heads = hv[3]; s.kv_heads = hv[4]; s.vocab = hv[5];
    s.max_ctx = max_ctx > 0 ? max_ctx : hv[6];
    s.weight_dtype = weight_dtype; s.kv_dtype = kv_dtype; s.lm_head_dtype = DT_F16;
    s.rope_order = rope_order >= 0 ? rope_order : 0;   // llama2.c stores q/k for interleaved rotation (qk_column_order 0)
    s.rope_theta = 10000.0f;
    s.layer_begin = layer_begin; s.layer_end = layer_end;
    s.weight_repack = 1;           // 3.5-bit files are kept as nibble blocks with the same values (see ifb_model_spec_t) s.norm_eps = 1e-5f; s.kq_scale = 1.0f; s.tp_rank = 0; s.tp_size = 1;
    ifb_model_t *m = nullptr;
    TRY(ifb_model_create(&m, &s));
    FILE *f = fopen(path, "rb");
    if (!f) { ifb_model_destroy(m); set_last_error("llama2.c: cannot open %s", path); return IFB_ERR_IO; }
    L2cHeader h;
    int rc = l2c_read_header(f, h);
    const int64_t d = s.d_model, hd = d / s.heads, kvd = (int64_t)s.kv_heads * hd, ff = s.d_ffn;
    std::vector<float> buf;
    auto load = [&](const char *name, int64_t rows, int64_t cols) -> int {
        buf.resize((size_t)rows * cols);
        if (fread(buf.data(), 4, buf.size(), f) != buf.size()) { set_last_error("llama2.c: %s is truncated at %s", path, name); return IFB_ERR_IO; }
        return ifb_model_load_tensor(m, name, buf.data(), DT_F32, rows, cols, stream);
    };
    if (rc == IFB_OK) rc = load("tok_embeddings", s.vocab, d);
    const bool keep_emb = rc == IFB_OK && shared;
    std::vector<float> emb;
    if (keep_emb) emb = buf;       // shared classifier: the embedding table is the lm_head too
    struct { const char *n; int64_t rows, cols; } per_layer[] = {
        {"attention_norm", 1, d}, {"wq", d, d}, {"wk", kvd, d}, {"wv", kvd, d}, {"wo", d, d},
        {"ffn_norm", 1, d}, {"w1", ff, d}, {"w2", d, ff}, {"w3", ff, d}};
    for (auto &t : per_layer)
        for (int l = 0; l < s.layers && rc == IFB_OK; l++) {
            char name[64];
            snprintf(name, sizeof(name), "layers.%d.%s", l, t.n);
            rc = load(name, t.rows, t.cols);
        }
    if (rc == IFB_OK) rc = load("output_norm", 1, d);
    if (rc == IFB_OK) {
        if (shared) {
            rc = ifb_model_load_tensor(m, "output", emb.data(), DT_F32, s.vocab, d, stream);
        } else {
            fseek(f, (long)((int64_t)h.seq * hd), SEEK_CUR);   // the reference skips exactly this many BYTES (model_reader.cc:3420)
            rc = load("output", s.vocab, d);
        }
    }
    fclose(f);
    if (rc != IFB_OK) { ifb_model_destroy(m); return rc; }
    *out = m;
    return IFB_OK;
}

// raw access for tests (read back weights / activations / cache)
extern "C" int32_t ifb_model_tensor_ptr(ifb_model_t *m, const char *name, void **ptr, int64_t *bytes) {
    IFB_CHECK_ARG(m && name && ptr && bytes, "tensor_ptr: null argument");
    std::string n(name);
    DevBuf *b = nullptr;
    if (n == "tok_embeddings") b = &m->tok_emb;
    else if (n == "output_norm") b = &m->out_norm;
    else if (n == "output") b = &m->lm_head;
    else if (n == "logits") b = &m->logits;
    else if (n == "x") b = &m->x0;
    else if (n == "x1") b = &m->x1;
    else if (n == "router" && m->spec.experts > 0) b = &m->router;
    else if (n == "moe_ids" && m->spec.experts > 0) b = &m->moe_ids;
    else if (n == "moe_wts" && m->spec.experts > 0) b = &m->moe_wts;
    else if (n == "hbuf") b = &m->hbuf;
    else if (n == "xchg" && m->spec.tp_size > 1) b = &m->xchg;
    else if (n == "qkv") b = &m->qkv;
    else if (n == "attn_out") b = &m->attn_out;
    else if (n == "scalars") b = &m->scalars;
    else if (n == "trace" && m->trace_on) b = &m->trace;
    else {
        int li = -1; char tn[64] = {0};
        if (sscanf(name, "layers.%d.%63s", &li, tn) == 2 && li >= m->l0 && li < m->l0 + m->nl) {
            LayerW &L = m->layers[li - m->l0];
            std::string t(tn);
            if (t == "attention_norm") b = &L.attn_norm; else if (t == "ffn_norm") b = &L.ffn_norm;
            else if (t == "wq") b = &L.wq; else if (t == "wk") b = &L.wk; else if (t == "wv") b = &L.wv;
            else if (t == "wo") b = &L.wo; else if (t == "w1") b = &L.w1; else if (t == "w2") b = &L.w2;
            else if (t == "w3") b = &L.w3; else if (t == "kcache") b = &L.kcache; else if (t == "vcache") b = &L.vcache;
            else if (t == "gate") b = &L.gate;
            else {
                int e = -1; char wn[16] = {0};
                if (sscanf(tn, "experts.%d.%15s", &e, wn) == 2 && e >= 0 && e < (int)L.ew1.size()) {
                    std::string w(wn);
                    if (w == "w1") b = &L.ew1[e]; else if (w == "w3") b = &L.ew3[e]; else if (w == "w2") b = &L.ew2[e];
                }
            }
            if (b != nullptr && b->p == nullptr) b = nullptr;   // e.g. "w1" of a MoE model
        }
    }
    if (!b) { set_last_error("tensor_ptr: unknown tensor '%s'", name); return IFB_ERR_ARG; }
    *ptr = b->p;
    *bytes = (int64_t)b->bytes;
    return IFB_OK;
}

// ---------------------------------------------------------------------------------------
// the decode step
// ---------------------------------------------------------------------------------------
__global__ void set_scalars_kernel(int32_t *sc, int pos, int token, const int32_t *token_dev) {
    sc[0] = pos;
    sc[1] = token_dev != nullptr ? *token_dev : token;
}
// next token + advance the position so the graph can be replayed back-to-back
// sc[3]: step sequence number, the flag base of the tensor-parallel exchange (every rank runs every step)
__global__ void advance_kernel(int32_t *sc, int advance) {
    if (advance) {
        sc[0] = sc[0] + 1;
        sc[1] = sc[2];
    }
    sc[3] = sc[3] + 1;
}

enum { PART_GEMV = 1, PART_ATTN = 2, PART_MISC = 4, PART_ALL = 7 };

// Build the plan of the persistent step kernel (ifb_step.cu) once the weights' addresses and, under tensor
// parallelism, the peers' exchange regions are known. Models it does not cover (mixture of experts, weight types
// without a dp4a path, head_dim other than 64/128) keep the multi-launch path: step_ok stays false.
static int step_prepare(ifb_model *m) {
    if (m->step_built) return IFB_OK;
    m->step_built = true;
    m->step_ok = false;
    const auto &s = m->spec;
    if (!m->use_step || s.experts > 0) return IFB_OK;
    StepParams &P = m->step;
    P = StepParams();
    P.n_layers = s.layers;
    P.d = s.d_model; P.q_dim = m->q_dim_l; P.kv_dim = m->kv_dim_l; P.ffn = m->ffn_l;
    P.heads = m->heads_l; P.kv_heads = m->kv_heads_l; P.hd = m->hd; P.group = m->heads_l / m->kv_heads_l;
    P.vocab = s.vocab; P.vocab_l = m->vocab_l;
    P.max_ctx = s.max_ctx; P.max_ctx_pad = (s.max_ctx + 63) / 64 * 64;
    P.rope_order = s.rope_order; P.rope_theta = s.rope_theta;
    P.w_dtype = s.weight_dtype;
    P.eps = s.norm_eps;
    P.softmax_scale = 1.0f / sqrtf((float)m->hd);     // alpha * kq_scale (inference_worker.cc:1666), as attn_decode
    {
        // warp-private pipelines (every consumer warp streams its own tasks) unless IFB_STEP_PRIV=0 or the model's
        // chunks do not fit: step_build then falls back to the shared-ring layout and clears P.priv
        const char *e = getenv("IFB_STEP_PRIV");
        P.priv = e ? (atoi(e) != 0) : 0;
    }
    if (step_build(P, s.kv_dtype, s.lm_head_dtype, sm_count()) != IFB_OK) return IFB_OK;   // not covered: reason in ifb_last_error
    if (s.rope_order != 0 && s.rope_order != 2) return IFB_OK;
    if (s.layers > 126 || s.d_model > 11520) return IFB_OK;   // word tags are 8 bits; a normalised vector is staged in one round
    std::vector<StepLayer> hl(s.layers);
    for (int l = 0; l < s.layers; l++) {
        const LayerW &L = m->layers[l];
        hl[l].wq = (const uint8_t *)L.wq.p; hl[l].wk = (const uint8_t *)L.wk.p; hl[l].wv = (const uint8_t *)L.wv.p;
        hl[l].wo = (const uint8_t *)L.wo.p; hl[l].w1 = (const uint8_t *)L.w1.p; hl[l].w3 = (const uint8_t *)L.w3.p;
        hl[l].w2 = (const uint8_t *)L.w2.p;
        hl[l].attn_norm = (const __half *)L.attn_norm.p; hl[l].ffn_norm = (const __half *)L.ffn_norm.p;
        hl[l].kcache = (uint8_t *)L.kcache.p; hl[l].vcache = (uint8_t *)L.vcache.p;
    }
    if (m->step_layers.p == nullptr) TRY(dev_alloc(m, m->step_layers, s.layers * sizeof(StepLayer)));
    IFB_CUDA(cudaMemcpy(m->step_layers.p, hl.data(), s.layers * sizeof(StepLayer), cudaMemcpyHostToDevice));
    if (m->step_ctr.p == nullptr) TRY(dev_alloc(m, m->step_ctr, (size_t)step_counter_bytes(P)));
    if (m->step_err == nullptr) {
        IFB_CUDA(cudaHostAlloc((void **)&m->step_err, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(m->step_err, 0, 64);
    }
    P.layers = (const StepLayer *)m->step_layers.p;
    P.tok_emb = (const __half *)m->tok_emb.p; P.out_norm = (const __half *)m->out_norm.p;
    P.lm_head = (const uint8_t *)m->lm_head.p;
    P.qkv = (__half *)m->qkv.p; P.attn_out = (__half *)m->attn_out.p; P.hbuf = (__half *)m->hbuf.p;
    P.ybuf = (__half *)m->ybuf.p; P.logits = (__half *)m->logits.p; P.x_out = (__half *)m->x0.p;
    P.partials = (float *)m->partials.p;
    P.qkvw = (uint2 *)m->qkvw.p; P.attw = (uint2 *)m->attw.p; P.hbufw = (uint2 *)m->hbufw.p;
    P.scalars = (int32_t *)m->scalars.p;
    P.cursors = (int *)m->step_ctr.p;
    P.done = (unsigned int *)((int *)m->step_ctr.p + P.n_ops);
    P.amax = (unsigned long long *)((uint8_t *)m->step_ctr.p + ((size_t)(2 * P.n_ops + s.layers + 1) * 4 + 7) / 8 * 8);
    P.parts = STEP_PART_ALL; P.advance = 1;
    P.tp_size = s.tp_size; P.tp_rank = s.tp_rank;
    if (s.tp_size > 1) {
        if (m->peer_xchg.empty()) { m->step_built = false; return IFB_OK; }   // not connected yet: decide at the first decode
        for (int p = 0; p < s.tp_size; p++) P.peer[p] = m->peer_xchg[p];
    } else {
        P.peer[0] = m->xchg.p;
        const char *e = getenv("IFB_STEP_LOCAL_LL");
        P.local_ll = e ? (atoi(e) ? 1 : 0) : 1;
    }
    P.x_recv = m->x_recv; P.x_recv_parity = m->x_recv_parity; P.x_recv_rank = m->x_recv_rank; P.x_logits = m->x_logits;
    void *dev_err = nullptr;
    IFB_CUDA(cudaHostGetDevicePointer(&dev_err, m->step_err, 0));
    P.err = (int *)dev_err;
    {
        const char *e = getenv("IFB_STEP_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 2000;
        P.timeout_ns = (unsigned long long)(ms > 0 ? ms : 2000) * 1000000ull;
    }
    {
        const char *e = getenv("IFB_STEP_BATCH");
        P.claim_batch = e ? atoi(e) : 4;
        if (P.claim_batch < 1) P.claim_batch = 1;
        const char *dbg = getenv("IFB_STEP_DEBUG");
        P.debug = dbg ? atoi(dbg) : 0;
        const char *f = getenv("IFB_STEP_INFLIGHT");
        P.max_inflight = f ? atoi(f) : 0;
        if (P.max_inflight >= P.nslots || P.max_inflight < 0) P.max_inflight = 0;
    }
    if (m->trace_on) {
        // [n_ops][grid][4] stamps; the multi-launch trace buffer is large enough for models up to ~800 ops
        if ((size_t)P.n_ops * sm_count() * 8 * 8 <= m->trace.bytes) P.trace = (unsigned long long *)m->trace.p;
    }
    m->step_ok = true;
    return IFB_OK;
}

static int issue_step_kernel(ifb_model *m, cudaStream_t st, int *launches, int parts) {
    StepParams P = m->step;
    P.parts = ((parts & PART_GEMV) ? STEP_PART_GEMV : 0) | ((parts & PART_ATTN) ? STEP_PART_ATT : 0);
    P.advance = parts == PART_ALL ? 1 : 0;
    IFB_CUDA(cudaMemsetAsync(m->step_ctr.p, 0, m->step_ctr.bytes, st));
    if (P.trace != nullptr) IFB_CUDA(cudaMemsetAsync(m->trace.p, 0, (size_t)P.n_ops * sm_count() * 8 * 8, st));
    TRY(step_launch(P, sm_count(), st));
    if (launches) *launches = 1;
    return IFB_OK;
}

static int issue_step(ifb_model *m, cudaStream_t st, int *launches, int parts = PART_ALL) {
    TRY(step_prepare(m));
    if (m->step_ok) return issue_step_kernel(m, st, launches, parts);
    const auto &s = m->spec;
    const int d = s.d_model, wdt = s.weight_dtype;
    const int pdl = m->use_pdl;
    int32_t *sc = (int32_t *)m->scalars.p;
    int n = 0;
    IFB_CUDA(cudaMemsetAsync(m->counters.p, 0, m->counters.bytes, st));
    auto CT = [&](int idx) -> void * { return (void *)((int *)m->counters.p + (size_t)idx * 8); };   // cursor + per-op completion counts
    auto TR = [&](int idx) -> void * { return (m->trace_on && idx < 512) ? (void *)((uint64_t *)m->trace.p + (size_t)idx * 160 * 8) : nullptr; };
    // a later pipeline stage starts from the hidden state its predecessor handed over (ifb_model_hidden_ptr(m, 0))
    if ((parts & PART_MISC) && m->first_stage) { TRY(embed_row(m->x0.p, m->tok_emb.p, DT_F16, d, 0, sc + 1, st, s.vocab)); n++; }
    // (the input has a buffer of its own: x0 is also where the step leaves its output, and a step may run twice --
    // the warm-up before the graph capture)
    if ((parts & PART_MISC) && !m->first_stage) IFB_CUDA(cudaMemcpyAsync(m->x0.p, m->xin.p, (size_t)d * 2, cudaMemcpyDeviceToDevice, st));
    const bool q8 = true;   // enable_full_quant_gemv = true (inference_engine.cc:53-57)
    auto ax8_ok = [](int dt) {
        return dt == DT_Q8_B32T2 || dt == DT_Q6_B64T1 || dt == DT_Q5_B64T1 || dt == DT_Q4_B32T1A ||
               dt == DT_Q4_B32T1B || dt == DT_Q4_B64T1 || dt == DT_Q3H_B64T1;
    };
    const int quant_x = (q8 && ax8_ok(wdt)) ? 1 : 0;
    __half *qkv = (__half *)m->qkv.p;
    const int tp = s.tp_size;
    TpExchange tpx;   // common fields
    tpx.size = tp; tpx.rank = s.tp_rank;
    if (tp > 1) for (int p = 0; p < tp; p++) tpx.peer[p] = m->peer_xchg[p];
    tpx.seq = sc + 3;
    // exchange index: wo of layer l -> 2l+1, w2 of layer l -> 2l+2; buffers alternate with the index parity
    // (a rank can run at most one exchange ahead of its slowest peer, so two buffers are enough)
    auto tp_push_ar = [&](int idx) { TpExchange t = tpx; t.push = 1; t.push_idx = (uint32_t)idx;
                              t.dst_off = m->x_recv + (idx & 1) * m->x_recv_parity + (int64_t)s.tp_rank * m->x_recv_rank; return t; };
    auto tp_gather_ar = [&](int idx, const void *resid, void *xout) { TpExchange t = tpx; t.gather = 1; t.gather_idx = (uint32_t)idx;
                              t.src_off = m->x_recv + (idx & 1) * m->x_recv_parity; t.src_rank_stride = m->x_recv_rank;
                              t.resid = resid; t.xout = xout; return t; };
    TpExchange t_head;
    // every GEMV of the step first (descriptors only), then the launches
    const int nl = m->nl;
    std::vector<GemvLaunch> g_qkv(nl), g_wo(nl), g_glu(nl), g_w2(nl);
    std::vector<TpExchange> t_qkv(nl), t_wo(nl), t_glu(nl), t_w2(nl);
    for (int li = 0; li < nl; li++) {
        LayerW &L = m->layers[li];
        // --- attention: norm + quantize + wq|wk|wv
        GemvLaunch &g = g_qkv[li];
        g.w_dtype = wdt; g.nseg = 3;
        g.w[0] = L.wq.p; g.w[1] = L.wk.p; g.w[2] = L.wv.p;
        g.rows[0] = m->q_dim_l; g.rows[1] = m->kv_dim_l; g.rows[2] = m->kv_dim_l;
        g.cx = d; g.x = m->x0.p; g.norm_w = L.attn_norm.p; g.norm_eps = s.norm_eps;
        g.y = qkv; g.quantize_x = quant_x; g.pdl = pdl; g.smem_budget = m->smem_kb[0] * 1024;
        if (tp > 1 && li > 0) { t_qkv[li] = tp_gather_ar(2 * li, m->x1.p, m->x0.p); g.tp = &t_qkv[li]; }   // x0 = x1 + sum(w2 partials) of the previous layer
        // --- wo + residual
        GemvLaunch &o = g_wo[li];
        o.w_dtype = wdt; o.nseg = 1; o.w[0] = L.wo.p; o.rows[0] = d; o.cx = m->q_dim_l;
        o.x = m->attn_out.p; o.residual = m->x0.p; o.y = m->x1.p; o.quantize_x = quant_x; o.pdl = pdl; o.smem_budget = m->smem_kb[1] * 1024;
        if (tp > 1) { o.residual = nullptr; t_wo[li] = tp_push_ar(2 * li + 1); o.tp = &t_wo[li]; }             // row-parallel: partial sums go to the peers
        // --- FFN: norm + quantize + (w1, w3) + SiLU*mul  (MoE: the same two launches per active expert slot,
        //     weights picked on the device: w[0] / w_gate below are only the alignment stand-ins)
        const bool moe = s.experts > 0;
        GemvLaunch &f = g_glu[li];
        f.w_dtype = wdt; f.nseg = 1; f.w[0] = moe ? L.ew1[0].p : L.w1.p; f.w_gate = moe ? L.ew3[0].p : L.w3.p; f.rows[0] = m->ffn_l; f.cx = d;
        f.x = m->x1.p; f.norm_w = L.ffn_norm.p; f.norm_eps = s.norm_eps; f.y = m->hbuf.p;
        f.quantize_x = quant_x; f.pdl = pdl; f.smem_budget = m->smem_kb[2] * 1024;
        if (tp > 1) { t_glu[li] = tp_gather_ar(2 * li + 1, m->x0.p, m->x1.p); f.tp = &t_glu[li]; }               // x1 = x0 + sum(wo partials)
        // --- w2 + residual
        GemvLaunch &w2 = g_w2[li];
        w2.w_dtype = wdt; w2.nseg = 1; w2.w[0] = moe ? L.ew2[0].p : L.w2.p; w2.rows[0] = d; w2.cx = m->ffn_l;
        w2.x = m->hbuf.p; w2.residual = m->x1.p; w2.y = m->x0.p; w2.quantize_x = quant_x; w2.pdl = pdl; w2.smem_budget = m->smem_kb[3] * 1024;
        if (tp > 1) { w2.residual = nullptr; t_w2[li] = tp_push_ar(2 * li + 2); w2.tp = &t_w2[li]; }
    }
    auto launch = [&](GemvLaunch *op) -> int {
        op->trace = TR(n); op->counter = CT(n);
        if (parts & PART_GEMV) { TRY(gemv_launch(*op, st)); n++; }
        return IFB_OK;
    };
    for (int li = 0; li < nl; li++) {
        LayerW &L = m->layers[li];
        TRY(launch(&g_qkv[li]));
        // --- RoPE(q,k) + append K,V + attention over the cache: one kernel
        AttnDecode ad;
        ad.k_new = qkv + m->q_dim_l; ad.v_new = qkv + m->q_dim_l + m->kv_dim_l;
        ad.rope = 1; ad.rope_order = s.rope_order; ad.rope_dims = 0; ad.rope_theta = s.rope_theta;
        ad.kv_dtype = s.kv_dtype; ad.heads = m->heads_l; ad.kv_heads = m->kv_heads_l; ad.head_dim = m->hd;
        ad.max_ctx = s.max_ctx; ad.ctx_len = 0; ad.kq_scale = s.kq_scale; ad.out = m->attn_out.p; ad.q = qkv;
        ad.kcache = L.kcache.p; ad.vcache = L.vcache.p; ad.workspace = m->attn_ws.p; ad.pos_dev = sc; ad.pdl = pdl;
        ad.trace = TR(n);
        if (parts & PART_ATTN) { TRY(attn_decode(ad, st)); n++; }
        if (s.experts > 0) {
            // ProcessGpuLayer_Moe (inference_worker.cc:1924-2146) without the host round trip: router GEMV on the
            // normalised x1 -> softmax + top-k on the device -> per active slot (ascending expert id) GLU + w2 with
            // device-selected weights -> out = sum_k w_k * y_k in half, + residual
            TRY(launch(&g_wo[li]));
            GemvLaunch r;
            r.w_dtype = DT_F16; r.nseg = 1; r.w[0] = L.gate.p; r.rows[0] = s.experts; r.cx = d;
            r.x = m->x1.p; r.norm_w = L.ffn_norm.p; r.norm_eps = s.norm_eps; r.y = m->router.p; r.quantize_x = 0; r.pdl = pdl;
            // tensor parallel: the router is the first reader of x1 = x0 + sum(wo partials) -- it gathers (and leaves x1
            // for the expert launches); the gate is duplicated, so every rank routes alike
            if (tp > 1) r.tp = &t_glu[li];
            TRY(launch(&r));
            if (parts & PART_GEMV) {
                TRY(moe_route((int32_t *)m->moe_ids.p, m->moe_wts.p, m->router.p, s.experts, s.experts_top_k,
                              s.moe_norm_top_k_prob, st));
                n++;
            }
            for (int k = 0; k < s.experts_top_k; k++) {
                GemvLaunch f = g_glu[li], w2 = g_w2[li];
                f.tp = nullptr; w2.tp = nullptr;        // x1 is in place; the slots' partial outputs stay local until the combine
                f.w_table = (const void *const *)L.tbl_w1.p; f.wg_table = (const void *const *)L.tbl_w3.p;
                f.moe_id = (const int32_t *)m->moe_ids.p + k;
                w2.w_table = (const void *const *)L.tbl_w2.p; w2.moe_id = (const int32_t *)m->moe_ids.p + k;
                w2.residual = nullptr; w2.y = (__half *)m->moe_y.p + (size_t)k * d;
                TRY(launch(&f));
                TRY(launch(&w2));
            }
            if (parts & PART_GEMV) {
                // tensor parallel: this rank's folded partial sums go to every rank (exchange 2l+2, where the dense w2 pushes);
                // the next layer's QKV GEMV (or the lm_head) adds them to x1
                TRY(moe_combine(m->x0.p, m->x1.p, m->moe_y.p, (const int32_t *)m->moe_ids.p, m->moe_wts.p,
                                s.experts_top_k, d, st, tp > 1 ? &t_w2[li] : nullptr));
                n++;
            }
        } else {
            TRY(launch(&g_wo[li]));
            TRY(launch(&g_glu[li]));
            TRY(launch(&g_w2[li]));
        }
    }
    if (!m->last_stage) {
        // the hidden state stays in x0 for the next stage; position += 1 so that decode_next works on every stage
        if (parts & PART_MISC) { advance_kernel<<<1, 1, 0, st>>>(sc, 1); n++; IFB_LAUNCH_CHECK(); }
        if (launches) *launches = n;
        return IFB_OK;
    }
    // --- final norm + lm_head (fp16 weights meet an fp16 activation: GemvHalf_AX_Alg3)
    GemvLaunch h;
    h.w_dtype = s.lm_head_dtype; h.nseg = 1; h.w[0] = m->lm_head.p; h.rows[0] = m->vocab_l; h.cx = d;
    h.x = m->x0.p; h.norm_w = m->out_norm.p; h.norm_eps = s.norm_eps; h.y = m->logits.p;
    h.quantize_x = (s.lm_head_dtype != DT_F16 && ax8_ok(s.lm_head_dtype)) ? 1 : 0; h.pdl = pdl;
    h.smem_budget = m->smem_kb[4] * 1024;
    h.trace = TR(n); h.counter = CT(n);
    if (tp > 1) {
        // input: x0 = x1 + sum(w2 partials of the last layer); output: this rank's logits slice to every rank
        t_head = tp_gather_ar(2 * s.layers, m->x1.p, m->x0.p);
        t_head.push = 1; t_head.push_idx = 255;
        t_head.dst_off = m->x_logits + (int64_t)s.tp_rank * m->vocab_l * 4;
        h.tp = &t_head;
    }
    if (parts & PART_GEMV) { TRY(gemv_launch(h, st)); n++; }
    if (parts & PART_MISC) {
        if (tp > 1) {
            TRY(argmax_tp(sc + 2, m->logits.p, (const uint8_t *)m->xchg.p + m->x_logits, s.vocab, sc + 3, 255, st));
        } else {
            TRY(argmax(sc + 2, m->logits.p, s.vocab, st));
        }
        n++;
        advance_kernel<<<1, 1, 0, st>>>(sc, 1); n++;
        IFB_LAUNCH_CHECK();
    } else if (tp > 1) {
        // a partial issue (measurement aid) is still one exchange round for the peers: new flags next time
        advance_kernel<<<1, 1, 0, st>>>(sc, 0);
        IFB_LAUNCH_CHECK();
    }
    if (launches) *launches = n;
    return IFB_OK;
}

// capture the step once (position and token live in device memory, so the graph replays without host data)
static int prepare_graph(ifb_model *m, cudaStream_t st) {
    if (!m->use_graph || m->graph_ready) return IFB_OK;
    {
        // warm every kernel once outside capture (cudaFuncSetAttribute etc. are not capturable)
        int32_t saved[3];   // pos, token, next token -- NOT the step sequence number (sc[3]): the warm-up step was a
                            // real exchange round for the peers
        IFB_CUDA(cudaStreamSynchronize(st));
        IFB_CUDA(cudaMemcpy(saved, m->scalars.p, sizeof(saved), cudaMemcpyDeviceToHost));
        TRY(issue_step(m, st, &m->launches_per_step));
        IFB_CUDA(cudaStreamSynchronize(st));
        IFB_CUDA(cudaMemcpy(m->scalars.p, saved, sizeof(saved), cudaMemcpyHostToDevice));
        cudaStream_t cs;
        IFB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        IFB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = issue_step(m, cs, &m->launches_per_step);
        cudaError_t e = cudaStreamEndCapture(cs, &m->graph);
        if (rc == IFB_OK && e == cudaSuccess) e = cudaGraphInstantiate(&m->graph_exec, m->graph, 0);
        cudaStreamDestroy(cs);
        if (rc != IFB_OK || e != cudaSuccess) {
            if (m->step_ok) {
                // a cooperative launch that this driver cannot capture: the step is one launch anyway, issue it directly
                cudaGetLastError();
                if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
                m->graph_exec = nullptr;
                m->use_graph = 0;
                return IFB_OK;
            }
            if (rc != IFB_OK) return rc;
            set_last_error("graph capture failed: %s", cudaGetErrorString(e));
            return IFB_ERR_CUDA;
        }
        m->graph_ready = true;
    }
    return IFB_OK;
}

extern "C" int32_t ifb_model_decode(ifb_model_t *m, int32_t token_id, const int32_t *token_dev, int32_t pos,
                                    ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr, "null model");
    IFB_CHECK_ARG(pos >= 0 && pos < m->spec.max_ctx, "decode: position %d outside max_context_len %d", pos, m->spec.max_ctx);
    IFB_CHECK_ARG(m->spec.tp_size == 1 || !m->peer_xchg.empty(), "decode: tensor-parallel model without imported peer buffers");
    IFB_CHECK_ARG(token_id < m->spec.vocab, "decode: token %d outside the vocabulary (%d)", token_id, m->spec.vocab);
    cudaStream_t st = (cudaStream_t)stream;
    if (pos >= 0 && (token_id >= 0 || token_dev != nullptr)) {
        set_scalars_kernel<<<1, 1, 0, st>>>((int32_t *)m->scalars.p, pos, token_id, token_id < 0 ? token_dev : nullptr);
        IFB_LAUNCH_CHECK();
    }
    m->host_pos = pos;
    if (!m->use_graph) return issue_step(m, st, &m->launches_per_step);
    TRY(prepare_graph(m, st));
    if (!m->use_graph) return issue_step(m, st, &m->launches_per_step);   // (a driver that cannot capture the launch)
    IFB_CUDA(cudaGraphLaunch(m->graph_exec, st));
    return IFB_OK;
}

// replay without touching the scalars: continues from the position/token the previous step left
extern "C" int32_t ifb_model_decode_next(ifb_model_t *m, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr && m->host_pos >= 0, "decode_next: call ifb_model_decode or ifb_model_prefill first");
    if (m->host_pos + 1 >= m->spec.max_ctx) {
        set_last_error("decode_next: position %d is outside max_context_len %d", m->host_pos + 1, m->spec.max_ctx);
        return IFB_ERR_STATE;
    }
    m->host_pos++;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(prepare_graph(m, st));        // (after a prefill the step has not been captured yet)
    if (!m->use_graph) return issue_step(m, st, &m->launches_per_step);
    IFB_CUDA(cudaGraphLaunch(m->graph_exec, st));
    return IFB_OK;
}

// measurement aid: issue only a subset of the step's kernels, eagerly (1 = the GEMV family,
// 2 = kv-append + attention, 4 = embedding/argmax/advance). Returns the number of launches (>= 0).
extern "C" int32_t ifb_model_issue_parts(ifb_model_t *m, int32_t parts, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr, "null model");
    int n = 0;
    int rc = issue_step(m, (cudaStream_t)stream, &n, parts);
    return rc != IFB_OK ? rc : n;
}

extern "C" const void *ifb_model_logits(const ifb_model_t *m) { return m ? m->logits.p : nullptr; }
extern "C" const int32_t *ifb_model_next_token(const ifb_model_t *m) { return m ? (const int32_t *)m->scalars.p + 2 : nullptr; }
extern "C" int32_t ifb_model_kernel_launches_per_step(const ifb_model_t *m) { return m ? m->launches_per_step : 0; }

// ---------------------------------------------------------------------------------------
// prompt pass: T tokens through every layer at once (rows > 1)
// ---------------------------------------------------------------------------------------
// The reference runs a prompt as ONE inference step with q_tokens = T (GpuInferenceWorker::Run, ProcessGpuLayer with
// rows > 1): RmsNorm rows, MatrixMultiplication rows > 1 = Dequantize + cuBLAS GEMM on the fp16 activations (no Q8
// activation quantization on this branch, inference_worker.cc:2364-2432), PositionEmbedding, SetK/VRows, batched attention,
// ... and the lm_head on the last row. Same here: ifb_gemm_q (tensor cores, weights dequantized tile by tile),
// kv_append (RoPE + cache rows), attn_prefill, elementwise kernels; then the decode machinery for the last row's logits.
static bool prefill_supported(const ifb_model *m) {
    const auto &s = m->spec;
    auto ok_w = [](int dt) { return dt == DT_F16 || dt == DT_Q8_B32T2 || dt == DT_Q6_B64T1 || dt == DT_Q5_B64T1 || dt == DT_Q4_B32T1A ||
                                    dt == DT_Q4_B32T1B || dt == DT_Q4_B64T1 || dt == DT_Q3H_B64T1; };
    return s.tp_size == 1 && s.experts == 0 && m->first_stage && m->last_stage && ok_w(s.weight_dtype) && (m->hd == 64 || m->hd == 128) && s.d_model % 128 == 0 &&
           m->q_dim_l % 128 == 0 && m->kv_dim_l % 128 == 0 && s.d_ffn % 128 == 0 && (s.rope_order == 0 || s.rope_order == 2);
}

// activations of up to n rows (prompt tokens / batched sequences), grown on demand
static int prefill_reserve(ifb_model *m, int n, cudaStream_t st) {
    if (n <= m->pf_cap) return IFB_OK;
    const auto &s = m->spec;
    const int d = s.d_model, qd = m->q_dim_l, kvd = m->kv_dim_l, f = s.d_ffn;
    IFB_CUDA(cudaStreamSynchronize(st));
    for (void **p : {&m->pf_x, &m->pf_xn, &m->pf_q, &m->pf_k, &m->pf_v, &m->pf_att, &m->pf_h1, &m->pf_h3, &m->pf_tmp, &m->pf_tok})
        if (*p) { cudaFree(*p); *p = nullptr; }
    const int cap = (n + 63) / 64 * 64;
    const size_t T = (size_t)cap;
    IFB_CUDA(cudaMalloc(&m->pf_x, T * d * 2));   IFB_CUDA(cudaMalloc(&m->pf_xn, T * d * 2));
    IFB_CUDA(cudaMalloc(&m->pf_q, T * qd * 2));  IFB_CUDA(cudaMalloc(&m->pf_k, T * kvd * 2)); IFB_CUDA(cudaMalloc(&m->pf_v, T * kvd * 2));
    IFB_CUDA(cudaMalloc(&m->pf_att, T * qd * 2)); IFB_CUDA(cudaMalloc(&m->pf_h1, T * f * 2)); IFB_CUDA(cudaMalloc(&m->pf_h3, T * f * 2));
    IFB_CUDA(cudaMalloc(&m->pf_tmp, T * d * 2)); IFB_CUDA(cudaMalloc(&m->pf_tok, T * 4));
    m->pf_cap = cap;
    return IFB_OK;
}

extern "C" int32_t ifb_model_prefill_supported(const ifb_model_t *m) { return (m != nullptr && prefill_supported(m)) ? 1 : 0; }

extern "C" int32_t ifb_model_prefill(ifb_model_t *m, const int32_t *tokens_host, int32_t n, int32_t pos0, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr && tokens_host != nullptr, "prefill: null argument");
    IFB_CHECK_ARG(n >= 1 && pos0 >= 0 && pos0 + n <= m->spec.max_ctx, "prefill: positions %d..%d outside max_context_len %d", pos0, pos0 + n, m->spec.max_ctx);
    IFB_CHECK_ARG(prefill_supported(m), "prefill: this model runs its prompt token by token (tensor parallel, mixture of experts, "
                  "a weight type without a tensor-core path, or dimensions that are not multiples of 128)");
    for (int i = 0; i < n; i++) IFB_CHECK_ARG(tokens_host[i] >= 0 && tokens_host[i] < m->spec.vocab, "prefill: token %d outside the vocabulary", tokens_host[i]);
    cudaStream_t st = (cudaStream_t)stream;
    const auto &s = m->spec;
    const int d = s.d_model, qd = m->q_dim_l, kvd = m->kv_dim_l, f = s.d_ffn, wdt = s.weight_dtype;
    TRY(prefill_reserve(m, n, st));
    IFB_CUDA(cudaMemcpyAsync(m->pf_tok, tokens_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    TRY(embed_rows(m->pf_x, m->tok_emb.p, (const int32_t *)m->pf_tok, n, d, s.vocab, st));
    for (int li = 0; li < s.layers; li++) {
        LayerW &L = m->layers[li];
        TRY(rmsnorm(m->pf_xn, m->pf_x, L.attn_norm.p, n, d, s.norm_eps, 0.f, st));
        TRY(gemm_q(wdt, m->pf_q, L.wq.p, m->pf_xn, n, qd, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_k, L.wk.p, m->pf_xn, n, kvd, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_v, L.wv.p, m->pf_xn, n, kvd, d, nullptr, st));
        KvAppend ka;
        ka.kv_dtype = s.kv_dtype; ka.kv_heads = m->kv_heads_l; ka.head_dim = m->hd; ka.max_ctx = s.max_ctx; ka.pos0 = pos0; ka.n_tokens = n;
        ka.kcache = L.kcache.p; ka.vcache = L.vcache.p; ka.k = m->pf_k; ka.v = m->pf_v;
        ka.rope = 1; ka.rope_order = s.rope_order; ka.rope_dims = 0; ka.rope_theta = s.rope_theta; ka.q = m->pf_q; ka.q_heads = m->heads_l;
        TRY(kv_append(ka, st));
        TRY(attn_prefill(s.kv_dtype, m->pf_att, m->pf_q, L.kcache.p, L.vcache.p, n, pos0, m->heads_l, m->kv_heads_l, m->hd, s.max_ctx, st));
        TRY(gemm_q(wdt, m->pf_tmp, L.wo.p, m->pf_att, n, d, qd, nullptr, st));
        TRY(add(m->pf_x, m->pf_x, m->pf_tmp, (int64_t)n * d, st));
        TRY(rmsnorm(m->pf_xn, m->pf_x, L.ffn_norm.p, n, d, s.norm_eps, 0.f, st));
        TRY(gemm_q(wdt, m->pf_h1, L.w1.p, m->pf_xn, n, f, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_h3, L.w3.p, m->pf_xn, n, f, d, nullptr, st));
        TRY(silu_mul(m->pf_h1, m->pf_h1, m->pf_h3, (int64_t)n * f, st));
        TRY(gemm_q(wdt, m->pf_tmp, L.w2.p, m->pf_h1, n, d, f, nullptr, st));
        TRY(add(m->pf_x, m->pf_x, m->pf_tmp, (int64_t)n * d, st));
    }
    // the last row: final norm + lm_head + greedy pick, and the device-side position / token for decode_next
    IFB_CUDA(cudaMemcpyAsync(m->x0.p, (const uint8_t *)m->pf_x + (size_t)(n - 1) * d * 2, (size_t)d * 2, cudaMemcpyDeviceToDevice, st));
    GemvLaunch h;
    h.w_dtype = s.lm_head_dtype; h.nseg = 1; h.w[0] = m->lm_head.p; h.rows[0] = m->vocab_l; h.cx = d;
    h.x = m->x0.p; h.norm_w = m->out_norm.p; h.norm_eps = s.norm_eps; h.y = m->logits.p;
    h.quantize_x = (s.lm_head_dtype != DT_F16) ? 1 : 0;
    TRY(gemv_launch(h, st));
    int32_t *sc = (int32_t *)m->scalars.p;
    set_scalars_kernel<<<1, 1, 0, st>>>(sc, pos0 + n - 1, tokens_host[n - 1], nullptr);
    IFB_LAUNCH_CHECK();
    TRY(argmax(sc + 2, m->logits.p, s.vocab, st));
    advance_kernel<<<1, 1, 0, st>>>(sc, 1);
    IFB_LAUNCH_CHECK();
    m->host_pos = pos0 + n - 1;
    return IFB_OK;
}

// ---------------------------------------------------------------------------------------
// batched decode: n sequences (each in its own KV slot, at its own position) advance by one token in ONE step.
// The reference batches the live queries of an engine the same way (max_concurrent_queries, query_state_table.h;
// inference_engine.cc:1054-1124: one Infer over all queries with q_tokens = sum of their new tokens): the matrix
// products see n rows -- here the dequant-fused tensor-core GEMM with split-K, the weights are read once for all
// sequences -- and attention runs per sequence over its own cache.
// ---------------------------------------------------------------------------------------
static bool batch_supported(const ifb_model *m) { return prefill_supported(m) && m->kv_slots > 1 && m->spec.vocab % 128 == 0; }

extern "C" int32_t ifb_model_batch_supported(const ifb_model_t *m) { return (m != nullptr && batch_supported(m)) ? 1 : 0; }
extern "C" void *ifb_model_batch_logits(ifb_model_t *m) { return m ? m->bt_logits.p : nullptr; }
extern "C" void *ifb_model_batch_next_tokens(ifb_model_t *m) { return m ? m->bt_next.p : nullptr; }

extern "C" int32_t ifb_model_decode_batch(ifb_model_t *m, const int32_t *tokens_host, const int32_t *slots_host,
                                          const int32_t *positions_host, int32_t n, ifb_stream_t stream) {
    IFB_CHECK_ARG(m != nullptr && tokens_host != nullptr && slots_host != nullptr && positions_host != nullptr, "decode_batch: null argument");
    IFB_CHECK_ARG(batch_supported(m), "decode_batch: needs kv_slots > 1 and a model the tensor-core path covers (single GPU, dense, "
                  "dimensions and vocabulary multiples of 128)");
    IFB_CHECK_ARG(n >= 1 && n <= m->kv_slots, "decode_batch: %d sequences, %d KV slots", n, m->kv_slots);
    unsigned seen = 0;
    for (int i = 0; i < n; i++) {
        IFB_CHECK_ARG(tokens_host[i] >= 0 && tokens_host[i] < m->spec.vocab, "decode_batch: token %d outside the vocabulary", tokens_host[i]);
        IFB_CHECK_ARG(slots_host[i] >= 0 && slots_host[i] < m->kv_slots && !((seen >> slots_host[i]) & 1u), "decode_batch: slot %d is out of range or named twice", slots_host[i]);
        IFB_CHECK_ARG(positions_host[i] >= 0 && positions_host[i] < m->spec.max_ctx, "decode_batch: position %d outside max_context_len %d", positions_host[i], m->spec.max_ctx);
        seen |= 1u << slots_host[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const auto &s = m->spec;
    const int d = s.d_model, qd = m->q_dim_l, kvd = m->kv_dim_l, f = s.d_ffn, wdt = s.weight_dtype;
    TRY(prefill_reserve(m, n, st));
    IFB_CUDA(cudaMemcpyAsync(m->pf_tok, tokens_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    IFB_CUDA(cudaMemcpyAsync(m->bt_pos.p, positions_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    TRY(embed_rows(m->pf_x, m->tok_emb.p, (const int32_t *)m->pf_tok, n, d, s.vocab, st));
    for (int li = 0; li < s.layers; li++) {
        LayerW &L = m->layers[li];
        TRY(rmsnorm(m->pf_xn, m->pf_x, L.attn_norm.p, n, d, s.norm_eps, 0.f, st));
        TRY(gemm_q(wdt, m->pf_q, L.wq.p, m->pf_xn, n, qd, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_k, L.wk.p, m->pf_xn, n, kvd, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_v, L.wv.p, m->pf_xn, n, kvd, d, nullptr, st));
        for (int b = 0; b < n; b++) {
            // RoPE(q, k) + append the new K/V row to the sequence's own cache + attention over it: the decode kernel
            AttnDecode ad;
            ad.q = (const __half *)m->pf_q + (size_t)b * qd;
            ad.k_new = (const __half *)m->pf_k + (size_t)b * kvd; ad.v_new = (const __half *)m->pf_v + (size_t)b * kvd;
            ad.rope = 1; ad.rope_order = s.rope_order; ad.rope_dims = 0; ad.rope_theta = s.rope_theta;
            ad.kv_dtype = s.kv_dtype; ad.heads = m->heads_l; ad.kv_heads = m->kv_heads_l; ad.head_dim = m->hd;
            ad.max_ctx = s.max_ctx; ad.ctx_len = 0; ad.kq_scale = s.kq_scale;
            ad.out = (__half *)m->pf_att + (size_t)b * qd;
            ad.kcache = (uint8_t *)L.kcache.p + (size_t)slots_host[b] * m->kv_slot_bytes;
            ad.vcache = (uint8_t *)L.vcache.p + (size_t)slots_host[b] * m->kv_slot_bytes;
            ad.workspace = m->attn_ws.p; ad.pos_dev = (const int32_t *)m->bt_pos.p + b; ad.pdl = 0;
            TRY(attn_decode(ad, st));
        }
        TRY(gemm_q(wdt, m->pf_tmp, L.wo.p, m->pf_att, n, d, qd, nullptr, st));
        TRY(add(m->pf_x, m->pf_x, m->pf_tmp, (int64_t)n * d, st));
        TRY(rmsnorm(m->pf_xn, m->pf_x, L.ffn_norm.p, n, d, s.norm_eps, 0.f, st));
        TRY(gemm_q(wdt, m->pf_h1, L.w1.p, m->pf_xn, n, f, d, nullptr, st));
        TRY(gemm_q(wdt, m->pf_h3, L.w3.p, m->pf_xn, n, f, d, nullptr, st));
        TRY(silu_mul(m->pf_h1, m->pf_h1, m->pf_h3, (int64_t)n * f, st));
        TRY(gemm_q(wdt, m->pf_tmp, L.w2.p, m->pf_h1, n, d, f, nullptr, st));
        TRY(add(m->pf_x, m->pf_x, m->pf_tmp, (int64_t)n * d, st));
    }
    // every row: final norm + lm_head + greedy pick
    TRY(rmsnorm(m->pf_xn, m->pf_x, m->out_norm.p, n, d, s.norm_eps, 0.f, st));
    TRY(gemm_q(s.lm_head_dtype, m->bt_logits.p, m->lm_head.p, m->pf_xn, n, s.vocab, d, nullptr, st));
    for (int b = 0; b < n; b++)
        TRY(argmax((int32_t *)m->bt_next.p + b, (const __half *)m->bt_logits.p + (size_t)b * s.vocab, s.vocab, st));
    return IFB_OK;
}

extern "C" int32_t ifb_model_tp_export(ifb_model_t *m, void *host_handle) {
    IFB_CHECK_ARG(m != nullptr && host_handle != nullptr, "tp_export: null argument");
    IFB_CHECK_ARG(m->spec.tp_size > 1 && m->xchg.p != nullptr, "tp_export: not a tensor-parallel model");
    cudaIpcMemHandle_t h;
    IFB_CUDA(cudaIpcGetMemHandle(&h, m->xchg.p));
    static_assert(sizeof(h) == 64, "IPC handle is 64 bytes");
    memcpy(host_handle, &h, sizeof(h));
    return IFB_OK;
}
extern "C" int32_t ifb_model_tp_import(ifb_model_t *m, const void *handles, int32_t n) {
    IFB_CHECK_ARG(m != nullptr && handles != nullptr, "tp_import: null argument");
    IFB_CHECK_ARG(n == m->spec.tp_size && n > 1, "tp_import: %d handles for a tensor-parallel group of %d", n, m->spec.tp_size);
    m->peer_xchg.assign(n, nullptr);
    for (int p = 0; p < n; p++) {
        if (p == m->spec.tp_rank) { m->peer_xchg[p] = m->xchg.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)p * 64, 64);
        void *ptr = nullptr;
        IFB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        m->peer_xchg[p] = ptr;
    }
    m->graph_ready = false;
    m->step_built = false;          // the step kernel's plan holds the peers' addresses
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
    return IFB_OK;
}

// The plan the persistent step kernel would run `spec` with, computed on the host alone (no device needed: sm_count
// falls back to 148): per op kind {units, tasks, R, cx, nblk, row_bytes, unit_bytes, vec4} and, for the warp-private
// schedule, {rt, nc, kc_blk, kc_bytes, ntasks}; then {priv, nslots, slot_bytes, smem_bytes, att_pos, pslot}. Lets the CPU
// tests check the geometry invariants (every row covered once, chunks within a slot, 16-byte bulk-copy granules).
extern "C" int32_t ifb_step_plan(const ifb_model_spec_t *spec, int32_t want_priv, int32_t *out, int32_t capacity) {
    IFB_CHECK_ARG(spec != nullptr && out != nullptr && capacity >= OPK_COUNT * 13 + 6, "step_plan: null argument or a buffer below %d ints", OPK_COUNT * 13 + 6);
    const ifb_model_spec_t &s = *spec;
    IFB_CHECK_ARG(s.heads > 0 && s.kv_heads > 0 && s.d_model % s.heads == 0, "step_plan: bad hyper-parameters");
    const int tp = s.tp_size > 0 ? s.tp_size : 1;
    StepParams P = StepParams();
    P.n_layers = s.layers;
    P.hd = s.d_model / s.heads;
    P.heads = s.heads / tp; P.kv_heads = s.kv_heads / tp; P.group = P.heads / (P.kv_heads > 0 ? P.kv_heads : 1);
    P.d = s.d_model; P.q_dim = P.heads * P.hd; P.kv_dim = P.kv_heads * P.hd; P.ffn = s.d_ffn / tp;
    P.vocab = s.vocab; P.vocab_l = s.vocab / tp;
    P.max_ctx = s.max_ctx; P.max_ctx_pad = (s.max_ctx + 63) / 64 * 64;
    P.w_dtype = (s.weight_repack && s.weight_dtype == DT_Q3H_B64T1) ? DT_Q4_B64T1 : s.weight_dtype;
    P.priv = want_priv ? 1 : 0;
    const int ldt = (s.weight_repack && s.lm_head_dtype == DT_Q3H_B64T1) ? DT_Q4_B64T1 : s.lm_head_dtype;
    const int rc = step_build(P, s.kv_dtype, ldt, 148);
    if (rc != IFB_OK) return rc;
    int k = 0;
    for (int kind = 0; kind < OPK_COUNT; kind++) {
        const StepOpShape &o = P.op[kind];
        const int v[8] = {o.units, o.tasks, o.R, o.cx, o.nblk, o.row_bytes, o.unit_bytes, o.vec4};
        for (int i = 0; i < 8; i++) out[k++] = v[i];
        const StepPrivShape &pv = P.pv[kind];
        const int w[5] = {pv.rt, pv.nc, pv.kc_blk, pv.kc_bytes, pv.ntasks};
        for (int i = 0; i < 5; i++) out[k++] = P.priv ? w[i] : 0;
    }
    out[k++] = P.priv; out[k++] = P.nslots; out[k++] = P.slot_bytes; out[k++] = P.smem_bytes; out[k++] = P.att_pos; out[k++] = P.pslot;
    return IFB_OK;
}

// 1 when the model decodes with the persistent step kernel, 0 for the multi-launch path; rec[4] (optional) receives
// the error record {code, op, detail, cta} a step kernel left before it trapped on a missed deadline (code 0: none)
extern "C" int32_t ifb_model_step_status(ifb_model_t *m, int32_t rec[4]) {
    if (m == nullptr) return 0;
    if (rec != nullptr) for (int i = 0; i < 4; i++) rec[i] = m->step_err ? ((volatile int *)m->step_err)[i] : 0;
    return m->step_ok ? 1 : 0;
}
