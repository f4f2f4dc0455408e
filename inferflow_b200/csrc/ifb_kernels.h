// Internal launch interface between the C ABI (ifb_capi.cu), the model runtime
// (ifb_model.cu) and the kernels. Not installed; the public surface is include/inferflow_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifb {

int sm_count();

// Tensor-parallel exchange through peer memory (NVLink): every rank owns one exchange region
// (same layout on all ranks) and holds the mapped base pointers of all peers.
//   push  : the GEMV epilogue writes its fp16 partial output into EVERY peer's region
//           (slot = this rank), then the last CTA publishes an epoch number to every peer;
//   gather: the next GEMV's prologue waits for all peers' epochs and reads
//           x = residual + sum over ranks (in rank order, fp16 adds like the reference's MergeTensors,
//           inference_worker.cc:2264-2289), CTA 0 stores x as the new residual stream.
// Tensor-parallel exchange through peer memory, flag-in-data ("LL") words: every 8-byte word of an exchange
// buffer is {two fp16 values, 32-bit flag}; a word is valid for an exchange when its flag equals
// (step sequence number << 8 | exchange index). 8-byte stores are atomic, so neither side needs a fence, a
// counter or a separate flag: the sender stores, the receiver polls the data words themselves.
struct TpExchange {
    int size = 1, rank = 0;
    void *peer[8] = {nullptr};     // exchange region of every rank as mapped in THIS process
    const int32_t *seq = nullptr;  // device: step sequence number (same value on every rank)
    // push (epilogue): this rank's outputs -> every rank's region
    int push = 0;
    int64_t dst_off = 0;           // byte offset (inside a region) of the word that receives output row 0
    uint32_t push_idx = 0;         // exchange index 1..255
    // gather (prologue): x = resid + sum over ranks of the partial vectors, in rank order
    int gather = 0;
    int64_t src_off = 0, src_rank_stride = 0;   // words of rank p's partial start at src_off + p*src_rank_stride
    uint32_t gather_idx = 0;
    const void *resid = nullptr;   // fp16 [cx]
    void *xout = nullptr;          // fp16 [cx] (written by CTA 0)
};

struct GemvLaunch {
    int w_dtype = 0;
    int nseg = 1;
    const void *w[3] = {nullptr, nullptr, nullptr};
    int rows[3] = {0, 0, 0};
    const void *w_gate = nullptr;
    int64_t cx = 0;
    const void *x = nullptr;       // fp16 [cx]
    const void *xq = nullptr;      // or Q8_B32T2 blocks [cx/32*34]
    const void *norm_w = nullptr;
    float norm_eps = 1e-5f, norm_base = 0.f;
    const void *residual = nullptr, *bias = nullptr;
    void *y = nullptr;
    int quantize_x = 1;
    int pdl = 0;
    int64_t smem_budget = 0;       // bytes of shared memory per CTA (0 = default: half an SM)
    void *trace = nullptr;         // optional timeline probe buffer (160*8 uint64)
    int l2_prefetch = 1;           // ask the L2 for the whole weight set at kernel entry (<= 64 MB)
    void *counter = nullptr;       // int32[8]: work cursor + per-op completion counts, must be 0 at launch (null: library scratch + memset)
    const TpExchange *tp = nullptr;
    // mixture of experts: the weights of THIS launch are w_table[*moe_id] (and wg_table[*moe_id] for the gate
    // matrix), the expert id being read on the device after the grid dependency; *moe_id < 0: the launch does nothing
    const void *const *w_table = nullptr, *const *wg_table = nullptr;
    const int32_t *moe_id = nullptr;
};
int gemv_launch(const GemvLaunch &g, cudaStream_t st);
// up to 4 GEMVs in ONE launch; op k+1 reads what op k wrote (counter: 8 ints, zero before the launch)

// rows > 1: Y[tokens][cy] = X[tokens][cx] . W[cy][cx]^T on the tensor cores, W dequantized tile by tile in shared memory
int gemm_q(int w_dtype, void *y_f16, const void *w, const void *x_f16, int64_t tokens, int64_t cy, int64_t cx,
           const void *bias_f16, cudaStream_t st);

// rows > 1 attention: T query tokens at positions pos0.. (K/V rows already appended) over cache positions 0..pos0+T-1
int attn_prefill(int kv_dtype, void *out_f16, const void *q_f16, const void *kcache, const void *vcache, int tokens, int pos0,
                 int heads, int kv_heads, int head_dim, int max_ctx, cudaStream_t st);
int embed_rows(void *dst_f16, const void *table_f16, const int32_t *tokens_dev, int n, int64_t cols, int vocab, cudaStream_t st);

int quantize_rows(int dt, void *dst, const void *src_f16, int64_t rows, int64_t cx, cudaStream_t st);
int dequantize_rows(int dt, void *dst_f16, const void *src, int64_t rows, int64_t cx, cudaStream_t st);
// Q3H_B64T1 blocks -> Q4_B64T1 blocks holding the same codes, base and scale (lossless; see ifb_quantize.cu)
int repack_q3h_q4b64(void *dst, const void *src, int64_t nblocks, cudaStream_t st);

int rmsnorm(void *y, const void *x, const void *w, int64_t rows, int64_t cols, float eps,
            float multi_base, cudaStream_t st);
int rope(void *x, int tokens, int heads, int head_dim, int rope_dims, int pos0, float theta,
         int order, cudaStream_t st);
int silu_mul(void *out, const void *a, const void *b, int64_t n, cudaStream_t st);
int add(void *c, const void *a, const void *b, int64_t n, cudaStream_t st);
int argmax(int32_t *out, const void *logits, int64_t n, cudaStream_t st);
// argmax after waiting until every peer has published `*epoch` in flags[0..n_peers)
int argmax_tp(int32_t *out, void *logits_f16, const void *ll_words, int64_t n, const int32_t *seq, uint32_t idx, cudaStream_t st);
int embed_row(void *dst_f16, const void *table, int table_dtype, int64_t cols, int token_id,
              const int32_t *token_dev, cudaStream_t st, int rows = 0);   // rows > 0: the token id is clamped to the table

// KV cache (head-major): per K or V buffer
//   F16      : half  [kv_heads][max_ctx][head_dim]
//   Q8_B32T2 : int8  [kv_heads][max_ctx][head_dim]  followed by  half scales [kv_heads][max_ctx][head_dim/32]
int64_t kv_cache_bytes(int kv_dtype, int kv_heads, int head_dim, int max_ctx);
struct KvAppend {
    int kv_dtype, kv_heads, head_dim, max_ctx, pos0, n_tokens;
    void *kcache, *vcache;
    const void *k, *v;       // fp16 [n][kv_heads*head_dim]
    // optional fused RoPE on K (and on q, in place) before the append
    int rope = 0, rope_order = 2, rope_dims = 0;
    float rope_theta = 10000.f;
    void *q = nullptr; int q_heads = 0;
    const int32_t *pos_dev = nullptr;   // position read from device memory (graph replay)
    int pdl = 0;
};
int kv_append(const KvAppend &a, cudaStream_t st);
int kv_get_rows(int kv_dtype, void *out, const void *cache, int kv_heads, int head_dim, int max_ctx,
                int pos0, int n, cudaStream_t st);
int64_t attn_workspace_bytes(int heads, int head_dim);
struct AttnDecode {
    int kv_dtype, heads, kv_heads, head_dim, max_ctx, ctx_len;
    float kq_scale;
    void *out; const void *q; const void *kcache, *vcache; void *workspace;
    const int32_t *pos_dev = nullptr;   // ctx_len = *pos_dev + 1 when set (graph replay)
    int pdl = 0;
    void *trace = nullptr;
    // fused append (layer runtime): the kernel itself applies RoPE to q (in shared memory only) and to the new
    // K row, quantizes and stores the new K/V row at position ctx_len-1, and attends over it -- replaces the
    // separate kv_append launch. k_new/v_new: fp16 [kv_heads][head_dim] of the current token.
    const void *k_new = nullptr, *v_new = nullptr;
    int rope = 0, rope_order = 2, rope_dims = 0;
    float rope_theta = 10000.0f;
};
int attn_decode(const AttnDecode &a, cudaStream_t st);

// router of one token: softmax over the experts' logits (SoftMax_Alg2 semantics), top-k (BuildRowsForMoE), slots in
// ascending expert order as ProcessGpuLayer_Moe walks them; ids[top_k] (-1 = unused), weights as half
int moe_route(int32_t *ids, void *weights_f16, const void *logits_f16, int experts, int top_k, int norm, cudaStream_t st);
// x_out = x_res + sum_k w_k * y_k with the reference's fp16 FMA chain (AddByRowIdx_Kernel, binary_tensor_opr.h:81)
// tp (optional, tp->push): y holds this rank's partial sums; the folded vector goes to every rank's exchange region
int moe_combine(void *x_out, const void *x_res, const void *y, const int32_t *ids, const void *weights_f16, int top_k,
                int64_t d, cudaStream_t st, const TpExchange *tp = nullptr);

int moe_topk(int32_t *ids, float *weights, int32_t *counts, const void *scores, int rows,
             int experts, int top_k, int norm, cudaStream_t st);

}  // namespace ifb
