/* inferflow_b200.h -- C ABI of the B200-native decode hot path of inferflow/inferflow.
 *
 * The reference has NO C-ABI / FFI seam: its operator layer is the
 * C++ static classes TensorMul (src/tensor/tensor_mul.h:21-57) and TensorOpr
 * (src/tensor/tensor_opr.h:39-130) called by GpuInferenceWorker
 * (src/transformer/inference_worker.cc). This header DEFINES the boundary a binding
 * would use; every entry point names the reference interface it replaces.
 * INTEGRATION.md shows the TensorMul/TensorOpr-shaped shim a maintainer would add.
 *
 * Conventions
 *  - all pointers are DEVICE pointers on the current CUDA device unless named host_*;
 *    the caller owns all memory; nothing here allocates or synchronises (the reference
 *    launchers end in cudaDeviceSynchronize, e.g. tensor_mul.cu:709 -- ours are async
 *    on `stream`, a cudaStream_t passed as void*).
 *  - element types are the reference's ElementType ids (src/tensor/tensor_common.h:15-42).
 *  - quantized tensors use the reference's block byte layouts unchanged
 *    (src/common/quant_types.h:11-174): rows of cx/capacity consecutive blocks.
 *  - return 0 on success, <0 on error (reference: LogError + return false);
 *    ifb_last_error() gives the message. There is no CPU fallback: without a CUDA
 *    device every compute entry returns IFB_E_CUDA.
 */
#ifndef INFERFLOW_B200_H
#define INFERFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFB_VERSION 100

#define IFB_E_ARG (-1)
#define IFB_E_TYPE (-2)
#define IFB_E_CUDA (-3)
#define IFB_E_STATE (-4)
#define IFB_E_IO (-5)

typedef void *ifb_stream_t; /* cudaStream_t */

/* ElementType (tensor_common.h:15-42) */
enum {
    IFB_F32 = 0, IFB_F16 = 1, IFB_BF16 = 2, IFB_I32 = 3, IFB_I16 = 4, IFB_Q8_GL = 5,
    IFB_Q8_LOG = 6, IFB_Q8_B32T1 = 7, IFB_Q8_B32T2 = 8, IFB_Q6_B64T1 = 9, IFB_Q5_B64T1 = 10,
    IFB_Q5_B32T1 = 11, IFB_Q4_B16 = 12, IFB_Q4_B32T1A = 13, IFB_Q4_B32T1B = 14,
    IFB_Q4_B32T2 = 15, IFB_Q4_B32P8 = 16, IFB_Q4_B64T1 = 17, IFB_Q3H_B64T1 = 18,
    IFB_Q3_B32T1A = 19, IFB_Q3_B32T1B = 20, IFB_Q2_B32T1A = 21, IFB_Q2_B32T1B = 22
};

/* ------------------------------------------------------------------ info */
int32_t ifb_version(void);
const char *ifb_last_error(void);
/* TensorCommon::ByteCount / block geometry (tensor_common.cc:162-167, quant_types.h) */
int32_t ifb_block_capacity(int32_t dtype);
int32_t ifb_block_bytes(int32_t dtype);
int64_t ifb_row_bytes(int32_t dtype, int64_t cx);
/* ini type names ("q4" -> Q4_B32T1A ...), tensor_common.cc:170-204; -1 if unknown */
int32_t ifb_dtype_from_name(const char *name);

/* plumbing for bindings that hold no CUDA runtime handle of their own (blocking copies) */
int32_t ifb_memcpy_d2h(void *host, const void *dev, int64_t bytes);
int32_t ifb_memcpy_h2d(void *dev, const void *host, int64_t bytes);
int32_t ifb_memcpy_d2d(void *dst_dev, const void *src_dev, int64_t bytes);   /* same device (a pipeline hand-off on one GPU) */
int32_t ifb_memcpy_d2h_async(void *host_pinned, const void *dev, int64_t bytes, ifb_stream_t stream);
int32_t ifb_memcpy_h2d_async(void *dev, const void *host_pinned, int64_t bytes, ifb_stream_t stream);
int32_t ifb_device_synchronize(void);

/* ------------------------------------------------------------------ codecs
 * TensorOpr::Quantize(B, A)  (tensor_opr.cu:1623; kernels/tensor_quant.h; the per-block
 * algorithms are Quantization::QuantizeRow_* in src/common/quantization.h:110-1194).
 * src fp16 [rows][cx] -> dst blocks [rows][row_bytes]. Codes are bit-exact with the
 * reference. Q8_B32T2 uses the device algorithm (tensor_quant.h:45-82) as the reference does. */
int32_t ifb_quantize_rows(int32_t dtype, void *dst, const void *src_f16, int64_t rows, int64_t cx,
                          ifb_stream_t stream);
/* TensorOpr::Dequantize(B, A) (tensor_opr.cu; kernels/tensor_quant.h:302-868) -> fp16 */
int32_t ifb_dequantize_rows(int32_t dtype, void *dst_f16, const void *src, int64_t rows, int64_t cx,
                            ifb_stream_t stream);

/* ------------------------------------------------------------------ GEMV
 * TensorMul::Gemv_AX(Y, A, X, Alg3, nullptr, bias) (tensor_mul.h:27, tensor_mul.cu:766):
 *   y[cy] = W[cy][cx] . x[cx] (+ bias), fp16 out.
 * x_dtype = IFB_Q8_B32T2 : dp4a path  (kernels/gemv.h:1500-1709; weights Q8_B32T2, Q6_B64T1,
 *                          Q5_B64T1, Q4_B32T1A/B, Q4_B64T1, Q3H_B64T1)
 * x_dtype = IFB_F16      : fp16-activation path (kernels/gemv.h:470-1470; any weight type
 *                          incl. IFB_F16), fp32 accumulation. */
int32_t ifb_gemv(int32_t w_dtype, void *y_f16, const void *w, int32_t x_dtype, const void *x,
                 int64_t cy, int64_t cx, const void *bias_f16, ifb_stream_t stream);

/* ------------------------------------------------------------------ GEMM (rows > 1: prefill, batched decode)
 * GpuInferenceWorker::MatrixMultiplication for more than one row (inference_worker.cc:2364-2432): the reference
 * dequantizes the WHOLE weight tensor to fp16 (:2375-2380), calls cublasGemmEx (:2413, cublas_engine.cu:425-505) and
 * transposes the result. Here:  Y[tokens][cy] = X[tokens][cx] . W[cy][cx]^T (+ bias), fp16 in / fp32 accumulate /
 * fp16 out, with W dequantized tile by tile in shared memory (value = fp16(q * scale + base), the bytes
 * TensorOpr::Dequantize writes) and multiplied on the 5th-generation tensor cores (tcgen05.mma, accumulator in tensor
 * memory). w_dtype: IFB_F16 or any block format of the dp4a family. cy % 128 == 0, cx % 64 == 0. */
int32_t ifb_gemm_q(int32_t w_dtype, void *y_f16, const void *w, const void *x_f16, int64_t tokens, int64_t cy,
                   int64_t cx, const void *bias_f16, ifb_stream_t stream);

/* Fused decode GEMV used by our layer runtime: one launch replaces
 *   [RmsNorm] -> Quantize(Q8_B32T2) -> 1..3 Gemv_AX sharing x -> [SiLU, Mul] -> [Add residual]
 * (inference_worker.cc:1038,1442-1444,1555-1571,1726-1900). */
typedef struct ifb_gemv_desc {
    int32_t w_dtype;
    int32_t nseg;            /* 1..3 weight matrices sharing x; outputs are concatenated in y */
    const void *w[3];
    int32_t rows[3];
    const void *w_gate;      /* optional (nseg==1): y = silu(w[0] x) * (w_gate x) */
    int64_t cx;
    const void *x_f16;       /* [cx] fp16 */
    const void *norm_w_f16;  /* optional fused RMSNorm weight [cx] (TensorOpr::RmsNorm, eps 1e-5) */
    float norm_eps;
    float norm_base;         /* multi_base */
    const void *residual_f16;/* optional [sum rows]: y += residual */
    const void *bias_f16;    /* optional [sum rows] */
    void *y_f16;             /* [sum rows] */
    int32_t quantize_x;      /* 1: x -> Q8_B32T2 then dp4a (reference default); 0: fp16 x */
    int32_t pdl;             /* 1: launch with programmatic stream serialization */
} ifb_gemv_desc_t;
int32_t ifb_gemv_fused(const ifb_gemv_desc_t *desc, ifb_stream_t stream);

/* ------------------------------------------------------------------ small ops
 * TensorOpr::RmsNorm (tensor_opr.cu:557; eps 1e-5 hard-coded at :568) */
int32_t ifb_rmsnorm(void *y_f16, const void *x_f16, const void *w_f16, int64_t rows, int64_t cols,
                    float eps, float multi_base, ifb_stream_t stream);
/* TensorOpr::PositionEmbedding ROPE (tensor_opr.cu:693-760): x [tokens][heads][head_dim] in place,
 * position of token t = pos0 + t; order 0 = interleaved (llama2.c), 2 = half-split (HF). */
int32_t ifb_rope(void *x_f16, int32_t tokens, int32_t heads, int32_t head_dim, int32_t rope_dims,
                 int32_t pos0, float theta, int32_t order, ifb_stream_t stream);
/* TensorOpr::Activation(SILU) + TensorOpr::Mul  (unary_tensor_opr.h:553, binary_tensor_opr.h) */
int32_t ifb_silu_mul(void *out_f16, const void *a_f16, const void *b_f16, int64_t n,
                     ifb_stream_t stream);
/* TensorOpr::Add */
int32_t ifb_add(void *c_f16, const void *a_f16, const void *b_f16, int64_t n, ifb_stream_t stream);
/* greedy token pick on device (replaces the D2H logits copy of SampleTokens,
 * inference_engine.cc:1986, for the Greedy strategy): first max wins */
int32_t ifb_argmax(int32_t *out_idx, const void *logits_f16, int64_t n, ifb_stream_t stream);

/* ------------------------------------------------------------------ KV cache + attention
 * GpuLayerKVCache (src/transformer/kv_cache.cc:104-249). Storage is OURS (head-major, codes
 * and scales separate) -- the values (int8 code, fp16 scale per 32 dims of a
 * row) are the reference's. kv_dtype = IFB_F16 or IFB_Q8_B32T2. head_dim % 32 == 0. */
int64_t ifb_kv_cache_bytes(int32_t kv_dtype, int32_t kv_heads, int32_t head_dim, int32_t max_ctx);
/* SetKRows+SetVRows: append n_tokens rows (fp16 [n][kv_heads*head_dim]) at position pos0 */
int32_t ifb_kv_append(int32_t kv_dtype, void *kcache, void *vcache, const void *k_f16,
                      const void *v_f16, int32_t kv_heads, int32_t head_dim, int32_t max_ctx,
                      int32_t pos0, int32_t n_tokens, ifb_stream_t stream);
/* GetKRows / GetVRows: rows pos0..pos0+n as fp16 [n][kv_heads*head_dim] (dequantized) */
int32_t ifb_kv_get_rows(int32_t kv_dtype, void *out_f16, const void *cache, int32_t kv_heads,
                        int32_t head_dim, int32_t max_ctx, int32_t pos0, int32_t n,
                        ifb_stream_t stream);
/* One decode step of grouped-query attention for one query token: replaces
 * CalculateProductKQ + SoftMax + Transpose/RepeatKV + Gemm(kqv) + Assign
 * (inference_worker.cc:1130-1320,1639-1724). q [heads][head_dim] fp16 (after RoPE);
 * ctx_len = positions 0..ctx_len-1 attended (the new token's K/V must already be appended).
 * workspace: ifb_attn_workspace_bytes(heads, head_dim) bytes. out [heads*head_dim] fp16. */
int64_t ifb_attn_workspace_bytes(int32_t heads, int32_t head_dim);
int32_t ifb_attn_decode(int32_t kv_dtype, void *out_f16, const void *q_f16, const void *kcache,
                        const void *vcache, int32_t heads, int32_t kv_heads, int32_t head_dim,
                        int32_t max_ctx, int32_t ctx_len, float kq_scale, void *workspace,
                        ifb_stream_t stream);

/* The same with the append fused in (what the layer runtime launches per layer and token): RoPE on q (kept in
 * shared memory, q_f16 is NOT modified) and on the new K row (PositionEmbedding, tensor_opr.cu:693-760), store of
 * the new K/V row at position ctx_len-1 (SetKRows/SetVRows, kv_cache.cc:104-183), then the attention above over
 * positions 0..ctx_len-1. k_new/v_new fp16 [kv_heads][head_dim]. rope_order < 0: no rotation. Results are
 * bit-identical to ifb_rope + ifb_kv_append + ifb_attn_decode. */
int32_t ifb_attn_decode_append(int32_t kv_dtype, void *out_f16, const void *q_f16, const void *k_new_f16,
                               const void *v_new_f16, void *kcache, void *vcache, int32_t heads,
                               int32_t kv_heads, int32_t head_dim, int32_t max_ctx, int32_t ctx_len,
                               float kq_scale, int32_t rope_order, int32_t rope_dims, float rope_theta,
                               void *workspace, ifb_stream_t stream);

/* Attention for q_tokens > 1 (the prompt): T query tokens at positions pos0..pos0+T-1, whose K/V rows are already in the
 * cache (ifb_kv_append), attend causally over positions 0..pos0+T-1. Replaces ProcessGpuLayer_Attention with rows > 1
 * (inference_worker.cc:1130-1320, :1639-1724: dequantize prefix, RepeatKV, Gemm_Alg2, SoftMax, Gemm_Alg2) without the
 * [heads][T][ctx] score tensor -- any prompt the cache holds works (the reference refuses beyond its scratch size,
 * :2612-2618). q / out fp16 [T][heads*head_dim], q with RoPE applied. head_dim 64 or 128. */
int32_t ifb_attn_prefill(int32_t kv_dtype, void *out_f16, const void *q_f16, const void *kcache, const void *vcache,
                         int32_t tokens, int32_t pos0, int32_t heads, int32_t kv_heads, int32_t head_dim,
                         int32_t max_ctx, ifb_stream_t stream);

/* ------------------------------------------------------------------ MoE router
 * HostTensorOpr::BuildRowsForMoE (src/tensor/host_tensor_opr.cc:190-244) on device:
 * scores fp16 [rows][experts] (softmax over all experts already applied, as in
 * ProcessGpuLayer_Moe) -> ids int32 [rows][top_k] (-1 = none), weights fp32, counts int32. */
int32_t ifb_moe_topk(int32_t *ids, float *weights, int32_t *counts, const void *scores_f16,
                     int32_t rows, int32_t experts, int32_t top_k, int32_t norm_top_k_prob,
                     ifb_stream_t stream);

/* The whole routing of ONE token on the device (no D2H): Tensor_SoftMax_Alg2_Kernel over the experts' logits
 * (unary_tensor_opr.h:481: fp32 max/exp/sum, exponentials stored as half, scaled by half(1/sum) in half), then
 * BuildRowsForMoE, slots in ascending expert id -- the order ProcessGpuLayer_Moe runs the experts in
 * (inference_worker.cc:2027-2040). logits fp16 [experts <= 32] -> ids int32 [top_k] (-1 = unused slot),
 * weights fp16 [top_k]. */
int32_t ifb_moe_route(int32_t *ids, void *weights_f16, const void *logits_f16, int32_t experts, int32_t top_k,
                      int32_t norm_top_k_prob, ifb_stream_t stream);
/* x_out[d] = x_res + sum over the used slots of w_k * y_k[d], accumulated as AddByRowIdx_Kernel does
 * (binary_tensor_opr.h:81: B = __hfma(A, w, B) in half, slots in order), then the layer's residual Add.
 * y fp16 [top_k][d]. */
int32_t ifb_moe_combine(void *x_out_f16, const void *x_res_f16, const void *y_f16, const int32_t *ids,
                        const void *weights_f16, int32_t top_k, int64_t d, ifb_stream_t stream);

/* ------------------------------------------------------------------ model runtime
 * The per-GPU decode runtime behind GpuInferenceWorker::Run / ProcessGpuLayer
 * (inference_worker.cc:234-977) for pre-norm llama-family decoders, one instance per GPU
 * (= per process; ranks of one tensor-parallel group exchange the row-parallel sums through
 * peer memory, see ifb_model_tp_*). */
typedef struct ifb_model_spec {
    int32_t layers, d_model, d_ffn, heads, kv_heads, vocab, max_ctx;
    int32_t weight_dtype;     /* device_weight_data_type */
    int32_t kv_dtype;         /* device_kv_cache_data_type: IFB_F16 or IFB_Q8_B32T2 */
    int32_t lm_head_dtype;    /* IFB_F16 (L > 20, network_builder.cc:829-844) or a quant type */
    int32_t rope_order;       /* qk_column_order: 0 interleaved, 2 half-split */
    float rope_theta;
    float norm_eps;
    float kq_scale;
    int32_t tp_rank, tp_size; /* tensor-parallel shard this instance holds (BY_COL/BY_ROW rules,
                                 network_builder.cc:1373-1406) */
    int32_t experts, experts_top_k; /* 0 = dense FFN; else every layer's FFN is a mixture of `experts` GLU FFNs,
                                       `experts_top_k` active per token (ProcessGpuLayer_Moe, inference_worker.cc:1924-2146) */
    int32_t moe_norm_top_k_prob;    /* hparams.moe_norm_top_k_prob (model.h:66): renormalise the selected weights */
    int32_t weight_repack;          /* 1: IFB_Q3H_B64T1 matrices are stored as IFB_Q4_B64T1 blocks holding the same codes,
                                       base and scale (lossless; 36 instead of 32 bytes per 64 weights, decoded with an
                                       order of magnitude fewer instructions). 0: stored as the reference stores them. */
    int32_t kv_slots;               /* KV caches held side by side (<= 16; 0 / 1 = one): the sequences ifb_model_decode_batch
                                       advances together (the reference's concurrent queries, query_state_table.h) */
    int32_t layer_begin, layer_end; /* pipeline stage (BY_LAYER / HYBRID, SplitGpuLayers): this instance holds layers
                                       [layer_begin, layer_end) only; 0, 0 = all. A stage with layer_begin > 0 takes
                                       its input from ifb_model_hidden_ptr(m, 0) instead of the embedding; one with
                                       layer_end < layers leaves its output in ifb_model_hidden_ptr(m, 1) and runs no lm_head. */
} ifb_model_spec_t;

typedef struct ifb_model ifb_model_t;

/* allocate weights / KV / activations for `spec` on the current device */
int32_t ifb_model_create(ifb_model_t **out, const ifb_model_spec_t *spec);
void ifb_model_destroy(ifb_model_t *m);
int64_t ifb_model_weight_bytes(const ifb_model_t *m);   /* quantized bytes read per token */
/* element type the layer matrices (which = 0) / the lm_head (which = 1) are STORED in: the spec's type unless
 * weight_repack changed it */
int32_t ifb_model_storage_dtype(const ifb_model_t *m, int32_t which);
/* the fp16 hidden state [d_model] of a pipeline stage: which = 0 the buffer a stage with layer_begin > 0 reads its input
 * from, which = 1 the buffer a stage leaves its output in. The caller hands it from stage to stage after the producing
 * step has completed (same GPU: ifb_memcpy_d2d; another GPU: cudaMemcpyPeer or a peer-mapped store). */
void *ifb_model_hidden_ptr(ifb_model_t *m, int32_t which);
int64_t ifb_model_kv_bytes_per_token(const ifb_model_t *m, int32_t ctx_len);
/* deterministic synthetic weights generated on the device: quant blocks
 * with uniform codes and |N(0,sigma)| fp16 scales, norms 1.0 */
int32_t ifb_model_init_synthetic(ifb_model_t *m, uint64_t seed, ifb_stream_t stream);
/* load one tensor from HOST fp32/fp16 data (DeviceTensorBuilder::Build, device_tensor_builder.cu:
 * 161-330): quantize on the device to the model's weight type and keep this rank's shard.
 * name: "tok_embeddings", "output_norm", "output", "layers.N.{attention_norm,wq,wk,wv,wo,
 * ffn_norm,w1,w2,w3}" ; data is the FULL (unsharded) [rows][cols] matrix */
int32_t ifb_model_load_tensor(ifb_model_t *m, const char *name, const void *host_data,
                              int32_t host_dtype, int64_t rows, int64_t cols, ifb_stream_t stream);
/* Read a model file in the llama2.c format the reference reads (ModelReader::LoadModel_Llama2DotC,
 * model_reader.cc:3245-3427: optional "ak42" v1 header or 7 x int32 {dim, hidden, layers, heads, kv_heads, +-vocab,
 * seq_len}, then fp32 tensors: tok_embeddings, all layers' attention_norm, wq, wk, wv, wo, ffn_norm, w1, w2, w3,
 * output_norm, seq_len*head_dim BYTES of rope tables (as the reference skips them), and the classifier when vocab < 0),
 * create the runtime for it (llama2.c RoPE order, fp16 lm_head, shared classifier when the file has none) and load
 * every tensor through ifb_model_load_tensor -- the same file drives the reference engine and this runtime.
 * header_out (optional): the 7 header ints as read (vocab made positive). */
int32_t ifb_llama2c_probe(const char *path, int32_t header_out[7], int32_t *shared_classifier);
int32_t ifb_model_load_llama2c(ifb_model_t **out, const char *path, int32_t weight_dtype, int32_t kv_dtype,
                               int32_t max_ctx, int32_t rope_order /* qk_column_order; < 0: the format's 0 */,
                               ifb_stream_t stream);
/* the same file, a pipeline stage of it: only layers [layer_begin, layer_end) are kept (0, 0 = all), the embedding on
 * the first stage, the final norm and the classifier on the last (BY_LAYER device groups, network_builder.cc:2094-2145) */
int32_t ifb_model_load_llama2c_stage(ifb_model_t **out, const char *path, int32_t weight_dtype, int32_t kv_dtype,
                                     int32_t max_context_len, int32_t qk_column_order, int32_t layer_begin,
                                     int32_t layer_end, ifb_stream_t stream);

/* The work plan of the persistent decode-step kernel for `spec`, computed on the host (no GPU needed): for each of the
 * 6 op kinds (wq|wk|wv, attention, wo, w1|w3, w2, lm_head) 13 ints {units, tasks, rows per unit, cx, blocks per row,
 * row bytes, unit bytes, vec4 | rows per task, K-chunks per row, blocks per chunk, bytes per chunk and row, tasks}
 * (the last five for the warp-private schedule, 0 otherwise), then {private schedule used, ring slots, slot bytes,
 * shared memory bytes, positions per attention unit, private slot bytes}. IFB_E_ARG when the kernel does not cover the
 * model (the runtime then decodes with the multi-launch path). Used by the CPU tests of the plan's invariants. */
int32_t ifb_step_plan(const ifb_model_spec_t *spec, int32_t want_private_schedule, int32_t *out, int32_t capacity);

/* Batched decode: n sequences, sequence i living in KV slot slots_host[i] at position positions_host[i], advance by
 * one token (tokens_host[i]) in one step -- the step InferenceEngine::Infer runs over all live queries
 * (inference_engine.cc:1054-1124). The matrix products see n rows (ifb_gemm_q: the weights are read once for all
 * sequences), attention runs per sequence over its own cache. Results on the device: fp16 logits [n][vocab] at
 * ifb_model_batch_logits, greedy next tokens int32 [n] at ifb_model_batch_next_tokens, rows in the order given. Needs
 * spec.kv_slots >= n and a model ifb_model_batch_supported accepts (single GPU, dense, dimensions and vocabulary
 * multiples of 128). Slot 0 is also the cache of ifb_model_decode / ifb_model_prefill. */
int32_t ifb_model_batch_supported(const ifb_model_t *m);
int32_t ifb_model_decode_batch(ifb_model_t *m, const int32_t *tokens_host, const int32_t *slots_host,
                               const int32_t *positions_host, int32_t n, ifb_stream_t stream);
void *ifb_model_batch_logits(ifb_model_t *m);
void *ifb_model_batch_next_tokens(ifb_model_t *m);

/* reset the KV length (and optionally fill `ctx_len` synthetic rows for benchmarking) */
int32_t ifb_model_kv_reset(ifb_model_t *m, int32_t ctx_len, int32_t fill_synthetic, uint64_t seed,
                           ifb_stream_t stream);
/* One decode step for one sequence at position pos (0-based): embedding row of `token_id`
 * (or *token_dev if token_id < 0) -> all layers -> logits (fp16 [vocab], device) and greedy
 * next token (int32, device). Captured as a CUDA graph internally; async on stream. */
int32_t ifb_model_decode(ifb_model_t *m, int32_t token_id, const int32_t *token_dev, int32_t pos,
                         ifb_stream_t stream);
/* The prompt in ONE pass (the reference's inference step with q_tokens = n): n tokens (host array) at positions
 * pos0..pos0+n-1 through every layer with the tensor-core GEMM and the prefill attention, K/V rows appended, logits and
 * greedy next token of the LAST prompt token left where ifb_model_decode leaves them; ifb_model_decode_next continues.
 * ifb_model_prefill_supported() == 0 (tensor parallel, MoE, dims not multiples of 128, weight types without a
 * tensor-core path): feed the prompt token by token with ifb_model_decode instead. */
int32_t ifb_model_prefill_supported(const ifb_model_t *m);
int32_t ifb_model_prefill(ifb_model_t *m, const int32_t *tokens_host, int32_t n, int32_t pos0, ifb_stream_t stream);
/* replay the step graph from the position / token the previous step left in device memory
 * (back-to-back generation without host round trips) */
int32_t ifb_model_decode_next(ifb_model_t *m, ifb_stream_t stream);
/* raw device pointer of a named tensor ("layers.N.wq", "logits", "layers.N.kcache", ...): tests
 * and tools read weights / activations back through it */
int32_t ifb_model_tensor_ptr(ifb_model_t *m, const char *name, void **ptr, int64_t *bytes);
/* measurement aid: issue only part of a step's kernels, eagerly (bit 0 = the GEMV family,
 * bit 1 = kv-append + attention, bit 2 = embedding/argmax/advance); returns #launches or <0 */
int32_t ifb_model_issue_parts(ifb_model_t *m, int32_t parts, ifb_stream_t stream);
const void *ifb_model_logits(const ifb_model_t *m);      /* device fp16 [vocab] */
const int32_t *ifb_model_next_token(const ifb_model_t *m); /* device int32 */
int32_t ifb_model_kernel_launches_per_step(const ifb_model_t *m);
/* 1 when the model decodes with the persistent step kernel (one launch per token; replaces GpuInferenceWorker::Run,
 * inference_worker.cc:234-977, and under tensor parallelism DistributeAndMergeTensors, :2148-2335), 0 when it keeps the
 * multi-launch path (mixture of experts, weight types without a dp4a kernel). rec (optional, 4 ints) receives the
 * record {code, op, detail, cta} a step kernel left in pinned host memory before it trapped on a missed deadline --
 * e.g. a tensor-parallel peer that never delivered its partial sums (code 0: none). Readable after a failed sync. */
int32_t ifb_model_step_status(ifb_model_t *m, int32_t rec[4]);
/* tensor-parallel plumbing: every rank exports the IPC handle of its exchange buffer,
 * the host side (torch.distributed all_gather / MPI / fork) hands all handles to every rank */
int32_t ifb_model_tp_export(ifb_model_t *m, void *host_handle_64B);
int32_t ifb_model_tp_import(ifb_model_t *m, const void *host_handles_64B_each, int32_t n);

/* ------------------------------------------------------------------ partitioner (host logic, no CUDA call)
 * The `devices` grammar of the engine config and what follows from it. `;` or `,` separate device GROUPS (layer /
 * pipeline split: contiguous layer ranges, the groups run one after the other), `&` or `|` join the devices INSIDE a
 * group (tensor split). Replaces InferenceEngine::LoadDeviceGroups (inference_engine.cc:1738-1783), the strategy rule
 * (inference_engine.cc:1509-1515), NetworkBuilder::SplitGpuLayers (network_builder.cc:2094-2117) and the tensor map
 * of Build{Attn,Ffn,Moe}TensorIdMap / AddLayerTasks_TensorParallel (network_builder.cc:1373-1413, 1594-1688). */
enum { IFB_STRATEGY_BY_LAYER = 0, IFB_STRATEGY_BY_TENSOR = 1, IFB_STRATEGY_HYBRID = 2 };   /* MultiGpuStrategy, network_structure.h:114 */
enum { IFB_PART_DUP = 0, IFB_PART_BY_ROW = 1, IFB_PART_BY_COL = 2 };                       /* TensorPartitionType, tensor_common.h:103 */
/* "0&1;2&3" -> ids_out = {0,1,2,3}, n_groups 2, group_size 2, strategy HYBRID. Ids are read with atoi, empty pieces
 * are skipped, all groups must have the same size (IFB_ERR_ARG and the reference's message otherwise). ids_out may be
 * NULL (sizes only). */
int32_t ifb_device_groups_parse(const char *devices, int32_t *ids_out, int32_t capacity, int32_t *n_groups,
                                int32_t *group_size, int32_t *strategy);
/* layers [start_layer, end_layer) over n_groups groups: ceil(L / n_groups) contiguous layers each, the last group takes
 * the rest, a group left without layers gets no range. ranges_out[2k], ranges_out[2k+1] = [begin, end) of range k,
 * group_of_range[k] = its group. (As in the reference the end of a range is (g+1)*per_group, not offset by
 * start_layer: identical for start_layer = 0, the only value the reference passes.) */
int32_t ifb_split_layers(int32_t start_layer, int32_t end_layer, int32_t n_groups, int32_t *ranges_out,
                         int32_t *group_of_range, int32_t *n_ranges);
/* how a layer tensor is divided inside a tensor-parallel group, by its name ("layers.3.wq", "w2", "ffn_norm" ...):
 * IFB_PART_BY_COL = output rows (wq wk wv qkv w1 w3 and their biases), IFB_PART_BY_ROW = contraction dimension
 * (wo w2), IFB_PART_DUP (norms, wo_b, w2_b, the MoE gate); -1 = not a layer tensor */
int32_t ifb_tensor_partition_type(const char *tensor_name);
/* element type a weight matrix gets on the device: weight_dtype, or IFB_F16 when its share per device
 * (elements / n_devices) is below tensor_quant_threshold (network_builder.cc:1558-1561, 1648-1652) */
int32_t ifb_tensor_device_dtype(int64_t elements, int32_t n_devices, int32_t weight_dtype, int64_t tensor_quant_threshold);
/* One process per GPU: rank r is the r-th device of the string. plan_out[IFB_PLAN_*] = its group, its rank and the size
 * of its tensor-parallel group, its layer range (ifb_model_spec.layer_begin / layer_end), the ranks it receives the
 * hidden state from / sends it to (-1: the embedding / the lm_head are its own), its device id, the strategy. */
enum { IFB_PLAN_GROUP = 0, IFB_PLAN_TP_RANK = 1, IFB_PLAN_TP_SIZE = 2, IFB_PLAN_LAYER_BEGIN = 3, IFB_PLAN_LAYER_END = 4,
       IFB_PLAN_PREV_RANK = 5, IFB_PLAN_NEXT_RANK = 6, IFB_PLAN_DEVICE = 7, IFB_PLAN_STRATEGY = 8, IFB_PLAN_FIELDS = 9 };
int32_t ifb_rank_plan(const char *devices, int32_t rank, int32_t n_layers, int32_t plan_out[IFB_PLAN_FIELDS]);

/* ------------------------------------------------------------------ engine-shaped host layer
 * The outermost seam of the reference for this path, for ONE live query and greedy decoding
 * (InferenceEngine, src/transformer/inference_engine.h:39-66,118): it reads the same .ini (layout of
 * bin/llm_inference*.ini: [transformer_engine] models / return_output_tensors, [model.<id>] model_dir,
 * model_specification_file, device_weight_data_type, device_kv_cache_data_type, max_context_len,
 * tensor_quant_threshold, [app_env.base] data_root_dir; ${config_dir} ${data_root_dir} ${model_name} macros) and the
 * model specification json, for model_file_format "llama2.c". Token ids in, token ids out (tokenizer and samplers
 * other than greedy are out of scope). */
typedef struct ifb_engine ifb_engine_t;
/* LoadConfig, resolved without touching the GPU. info_out = {weight_dtype, kv_dtype, max_context_len,
 * qk_column_order, return_output_tensors} */
int32_t ifb_engine_resolve(const char *ini_path, const char *section, const char *data_root_dir,
                           char *model_path_out, int32_t path_capacity, int32_t info_out[5]);
/* LoadConfig + Init */
int32_t ifb_engine_create(ifb_engine_t **out, const char *ini_path, const char *section, const char *data_root_dir);
void ifb_engine_destroy(ifb_engine_t *e);
int32_t ifb_engine_vocab_size(const ifb_engine_t *e);
/* AddQuery: > 0 query id, 0 busy, < 0 error */
int32_t ifb_engine_add_query(ifb_engine_t *e, const int32_t *tokens, int32_t n);
/* Infer: the whole prompt on the first call, the committed token afterwards; logits_f32 optional [vocab] */
int32_t ifb_engine_infer(ifb_engine_t *e, int32_t *next_token, float *logits_f32);
/* CommitInferenceResult / RemoveQuery */
int32_t ifb_engine_commit(ifb_engine_t *e, int32_t token, int32_t is_end);
int32_t ifb_engine_remove_query(ifb_engine_t *e, int32_t query_id);

#ifdef __cplusplus
}
#endif
#endif /* INFERFLOW_B200_H */
