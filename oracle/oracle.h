/* oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's algorithms on the decode hot path
 * (SURVEY.md section 8a). Each function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
 * may load this library; the product (inferflow_b200/) never does.
 *
 * Pinning: the reference ships no tests / golden vectors, so
 * this restatement is pinned against the reference ITSELF compiled here
 * (oracle/_ref/libifref_cpu.so, built by oracle/Makefile from /root/reference)
 * -- see tests/golden/make_golden.py and tests/test_oracle_vs_ref.py -- and
 * against the committed fixtures that script generated (tests/golden/*.npz).
 */
#ifndef IFB_ORACLE_H
#define IFB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type ids = reference ElementType (src/tensor/tensor_common.h:15-42) */
enum {
    ORC_F32 = 0, ORC_F16 = 1, ORC_Q8_B32T1 = 7, ORC_Q8_B32T2 = 8, ORC_Q6_B64T1 = 9,
    ORC_Q5_B64T1 = 10, ORC_Q5_B32T1 = 11, ORC_Q4_B16 = 12, ORC_Q4_B32T1A = 13,
    ORC_Q4_B32T1B = 14, ORC_Q4_B64T1 = 17, ORC_Q3H_B64T1 = 18, ORC_Q3_B32T1A = 19,
    ORC_Q3_B32T1B = 20, ORC_Q2_B32T1A = 21, ORC_Q2_B32T1B = 22
};

int orc_block_capacity(int dt);
int orc_block_bytes(int dt);
int64_t orc_row_bytes(int dt, int64_t cx);

/* fp16 helpers (IEEE round-to-nearest-even) */
uint16_t orc_f32_to_f16(float v);
float orc_f16_to_f32(uint16_t h);
void orc_round_f16(float *v, int64_t n); /* v[i] = (float)(half)v[i] */

/* Quantization::QuantizeRow_* (src/common/quantization.h:110-1194): src is fp32
 * (values the device sees as fp16 should already be fp16-representable).
 * Q8_B32T2 follows the DEVICE algorithm (kernels/tensor_quant.h:45-82, alg 2), which is
 * what TensorOpr::Quantize always selects (tensor_opr.cu:1645); `orc_quantize_row_q8t2_host`
 * is the host flavour (quantization.h:186-221). Returns 0 / -1 (unsupported type). */
int orc_quantize_row(int dt, uint8_t *dst, const float *src, int64_t n);
int orc_quantize_row_q8t2_host(uint8_t *dst, const float *src, int64_t n);
/* Quantization::Dequantize* : dst fp32 = q*scale+base (no fp16 rounding of the result) */
int orc_dequantize_row(int dt, float *dst, const uint8_t *src, int64_t n);
/* Quantization::GetInt4 for the dp4a path (codes of the block as int8/uint8) */
int orc_block_codes(int dt, int32_t *codes, float *scale, float *base, const uint8_t *blk);

/* GEMV y = W x, W [cy][cx] block-quantized rows.
 * ax8: x given as Q8_B32T2 blocks; math of kernels/gemv.h:1500-1709
 *      y = sum_b s_x[b] * (scale_w*sum(q_w q_x) + base_w*sum(q_x)), accumulated in double.
 * half: x fp16 values (as float); dequantized weights rounded to fp16 first
 *      (kernels/gemv.h:633-1470: `half arr_a[]`), products accumulated in double.
 * Output y is NOT rounded to fp16 (callers decide). */
int orc_gemv_ax8(int dt, double *y, const uint8_t *w, const uint8_t *xq, int64_t cy, int64_t cx);
int orc_gemv_half(int dt, double *y, const uint8_t *w, const float *x, int64_t cy, int64_t cx);
/* fp16 weights (kernels/gemv.h:470), math in double */
int orc_gemv_f16(double *y, const uint16_t *w, const float *x, int64_t cy, int64_t cx);

/* Tensor_RmsNorm_Kernel (kernels/unary_tensor_opr.h:217): y = x*rsqrt(mean(x^2)+eps)*(base+w) */
void orc_rmsnorm(float *y, const float *x, const float *w, int64_t n, float eps, float multi_base);

/* PosEmbedding_Rope_{Std,Order2}_Kernel (unary_tensor_opr.h:662,700) on one token:
 * x [heads][head_dim] in place; order 0 = interleaved pairs, 2 = half-split (HF). */
void orc_rope(float *x, int heads, int head_dim, int rope_dims, int pos, float theta, int order);

/* SiluActivation_Kernel (unary_tensor_opr.h:553) + Mul: out = silu(a)*b */
void orc_silu_mul(float *out, const float *a, const float *b, int64_t n);

/* Decode attention for ONE query token (GpuInferenceWorker::CalculateProductKQ +
 * SoftMax + kqv, inference_worker.cc:1130-1320,1639-1724; softmax kernel
 * unary_tensor_opr.h:481): q [heads][head_dim] (after RoPE), K,V [ctx][kv_heads*head_dim]
 * given as fp32 (already dequantized), GQA via head/heads_per_group.
 * exact==1: everything in double. exact==0: reproduce the reference's fp16 storage of
 * kq, of the exp values, of the normalised probabilities and of the output. */
void orc_attn_decode(float *out, const float *q, const float *k, const float *v, int ctx,
                     int heads, int kv_heads, int head_dim, float kq_scale, int exact);

/* HostTensorOpr::BuildRowsForMoE (src/tensor/host_tensor_opr.cc:190-244): scores
 * [rows][experts] -> top-k ids/weights per row. Returns number selected per row in cnt. */
void orc_moe_topk(int32_t *ids, float *weights, int32_t *cnt, const float *scores, int rows,
                  int experts, int top_k, int norm_top_k_prob, float min_score);

#ifdef __cplusplus
}
#endif
#endif
