// extern "C" entry points of include/inferflow_b200.h (operator level). The model runtime's
// entry points live in ifb_model.cu. Each wrapper validates like the reference launcher it
// replaces (shape / type errors -> message + negative code; the reference logs and returns
// false) and forwards to the kernels. No CPU fallback anywhere: a missing CUDA device
// surfaces as IFB_E_CUDA from the first launch.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <strings.h>

#include "ifb_formats.cuh"
#include "ifb_kernels.h"
#include "../../include/inferflow_b200.h"

namespace ifb {
static thread_local char g_err[512] = "";
void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ifb

using namespace ifb;

extern "C" {

int32_t ifb_version(void) { return IFB_VERSION; }
const char *ifb_last_error(void) { return g_err; }
int32_t ifb_block_capacity(int32_t dt) { return block_capacity(dt); }
int32_t ifb_block_bytes(int32_t dt) { return block_bytes(dt); }
int64_t ifb_row_bytes(int32_t dt, int64_t cx) { return row_bytes(dt, cx); }

// tensor_common.cc:170-204 (case-insensitive)
int32_t ifb_dtype_from_name(const char *name) {
    static const struct { const char *n; int dt; } tbl[] = {
        {"f32", DT_F32}, {"f16", DT_F16}, {"bf16", DT_BF16},
        {"q8_gl", DT_Q8_GL}, {"q8_log", DT_Q8_LOG},
        {"q8_b32t1", DT_Q8_B32T1}, {"q8_b32t2", DT_Q8_B32T2}, {"q8", DT_Q8_B32T2},
        {"q6_b64t1", DT_Q6_B64T1}, {"q6", DT_Q6_B64T1},
        {"q5_b32t1", DT_Q5_B32T1}, {"q5_b64t1", DT_Q5_B64T1}, {"q5", DT_Q5_B64T1},
        {"q4_b16", DT_Q4_B16}, {"q4_b32t1a", DT_Q4_B32T1A}, {"q4_b32t1b", DT_Q4_B32T1B},
        {"q4_b32t1", DT_Q4_B32T1A}, {"q4_b32t2", DT_Q4_B32T2}, {"q4_b32p8", DT_Q4_B32P8},
        {"q4_b64t1", DT_Q4_B64T1}, {"q4", DT_Q4_B32T1A},
        {"q3h_b64t1", DT_Q3H_B64T1}, {"q3h", DT_Q3H_B64T1},
        {"q3_b32t1a", DT_Q3_B32T1A}, {"q3_b32t1b", DT_Q3_B32T1B}, {"q3_b32t1", DT_Q3_B32T1B}, {"q3", DT_Q3_B32T1B},
        {"q2_b32t1a", DT_Q2_B32T1A}, {"q2_b32t1b", DT_Q2_B32T1B}, {"q2_b32t1", DT_Q2_B32T1B}, {"q2", DT_Q2_B32T1B},
    };
    if (!name) return -1;
    for (auto &e : tbl)
        if (strcasecmp(name, e.n) == 0) return e.dt;
    return -1;
}

// plumbing for bindings without their own CUDA runtime handle (tests, tools)
int32_t ifb_memcpy_d2h(void *host, const void *dev, int64_t bytes) {
    IFB_CUDA(cudaMemcpy(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost));
    return IFB_OK;
}
int32_t ifb_memcpy_h2d(void *dev, const void *host, int64_t bytes) {
    IFB_CUDA(cudaMemcpy(dev, host, (size_t)bytes, cudaMemcpyHostToDevice));
    return IFB_OK;
}
int32_t ifb_memcpy_d2d(void *dst, const void *src, int64_t bytes) {
    IFB_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice));
    return IFB_OK;
}
int32_t ifb_memcpy_d2h_async(void *host, const void *dev, int64_t bytes, ifb_stream_t st) {
    IFB_CUDA(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)st));
    return IFB_OK;
}
int32_t ifb_memcpy_h2d_async(void *dev, const void *host, int64_t bytes, ifb_stream_t st) {
    IFB_CUDA(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)st));
    return IFB_OK;
}
int32_t ifb_device_synchronize(void) {
    IFB_CUDA(cudaDeviceSynchronize());
    return IFB_OK;
}

int32_t ifb_quantize_rows(int32_t dt, void *dst, const void *src, int64_t rows, int64_t cx, ifb_stream_t st) {
    IFB_CHECK_ARG(dst && src, "Quantize: null tensor");
    return quantize_rows(dt, dst, src, rows, cx, (cudaStream_t)st);
}
int32_t ifb_dequantize_rows(int32_t dt, void *dst, const void *src, int64_t rows, int64_t cx, ifb_stream_t st) {
    IFB_CHECK_ARG(dst && src, "Dequantize: null tensor");
    return dequantize_rows(dt, dst, src, rows, cx, (cudaStream_t)st);
}

int32_t ifb_gemv(int32_t w_dtype, void *y, const void *w, int32_t x_dtype, const void *x, int64_t cy,
                 int64_t cx, const void *bias, ifb_stream_t st) {
    IFB_CHECK_ARG(y && w && x, "Gemv_AX: null tensor");
    IFB_CHECK_ARG(x_dtype == DT_F16 || x_dtype == DT_Q8_B32T2, "Gemv_AX: X must be F16 or Q8_B32T2 (got %d)", x_dtype);
    IFB_CHECK_ARG(cy > 0 && cy < (1ll << 31), "Gemv_AX: bad cy");
    GemvLaunch g;
    g.w_dtype = w_dtype; g.nseg = 1; g.w[0] = w; g.rows[0] = (int)cy; g.cx = cx;
    if (x_dtype == DT_Q8_B32T2) { g.xq = x; g.quantize_x = 1; } else { g.x = x; g.quantize_x = 0; }
    g.bias = bias; g.y = y;
    return gemv_launch(g, (cudaStream_t)st);
}

int32_t ifb_gemm_q(int32_t w_dtype, void *y, const void *w, const void *x, int64_t tokens, int64_t cy, int64_t cx,
                   const void *bias, ifb_stream_t st) {
    IFB_CHECK_ARG(y && w && x, "MatrixMultiplication: null tensor");
    IFB_CHECK_ARG(cy > 0 && cy < (1ll << 31) && cx > 0 && cx < (1ll << 31) && tokens > 0 && tokens < (1ll << 31), "MatrixMultiplication: bad shape");
    return gemm_q(w_dtype, y, w, x, tokens, cy, cx, bias, (cudaStream_t)st);
}

int32_t ifb_gemv_fused(const ifb_gemv_desc_t *d, ifb_stream_t st) {
    IFB_CHECK_ARG(d != nullptr, "gemv_fused: null descriptor");
    IFB_CHECK_ARG(d->x_f16 && d->y_f16, "gemv_fused: null tensor");
    GemvLaunch g;
    g.w_dtype = d->w_dtype; g.nseg = d->nseg;
    for (int i = 0; i < 3; i++) { g.w[i] = d->w[i]; g.rows[i] = d->rows[i]; }
    g.w_gate = d->w_gate; g.cx = d->cx; g.x = d->x_f16; g.norm_w = d->norm_w_f16;
    g.norm_eps = d->norm_eps; g.norm_base = d->norm_base; g.residual = d->residual_f16;
    g.bias = d->bias_f16; g.y = d->y_f16; g.quantize_x = d->quantize_x; g.pdl = d->pdl;
    return gemv_launch(g, (cudaStream_t)st);
}

int32_t ifb_rmsnorm(void *y, const void *x, const void *w, int64_t rows, int64_t cols, float eps,
                    float base, ifb_stream_t st) {
    IFB_CHECK_ARG(y && x, "RmsNorm: null tensor");
    return rmsnorm(y, x, w, rows, cols, eps, base, (cudaStream_t)st);
}
int32_t ifb_rope(void *x, int32_t tokens, int32_t heads, int32_t hd, int32_t rope_dims, int32_t pos0,
                 float theta, int32_t order, ifb_stream_t st) {
    IFB_CHECK_ARG(x, "PositionEmbedding: null tensor");
    return rope(x, tokens, heads, hd, rope_dims, pos0, theta, order, (cudaStream_t)st);
}
int32_t ifb_silu_mul(void *o, const void *a, const void *b, int64_t n, ifb_stream_t st) {
    IFB_CHECK_ARG(o && a && b, "Activation: null tensor");
    return silu_mul(o, a, b, n, (cudaStream_t)st);
}
int32_t ifb_add(void *c, const void *a, const void *b, int64_t n, ifb_stream_t st) {
    IFB_CHECK_ARG(c && a && b, "Add: null tensor");
    return add(c, a, b, n, (cudaStream_t)st);
}
int32_t ifb_argmax(int32_t *out, const void *logits, int64_t n, ifb_stream_t st) {
    IFB_CHECK_ARG(out && logits, "argmax: null tensor");
    return argmax(out, logits, n, (cudaStream_t)st);
}

int64_t ifb_kv_cache_bytes(int32_t kv_dtype, int32_t kv_heads, int32_t hd, int32_t max_ctx) {
    return kv_cache_bytes(kv_dtype, kv_heads, hd, max_ctx);
}
int32_t ifb_kv_append(int32_t kv_dtype, void *kc, void *vc, const void *k, const void *v, int32_t kv_heads,
                      int32_t hd, int32_t max_ctx, int32_t pos0, int32_t n, ifb_stream_t st) {
    IFB_CHECK_ARG(kc && vc && k && v, "SetKRows: null tensor");
    KvAppend a;
    a.kv_dtype = kv_dtype; a.kv_heads = kv_heads; a.head_dim = hd; a.max_ctx = max_ctx; a.pos0 = pos0;
    a.n_tokens = n; a.kcache = kc; a.vcache = vc; a.k = k; a.v = v;
    return kv_append(a, (cudaStream_t)st);
}
int32_t ifb_kv_get_rows(int32_t kv_dtype, void *out, const void *cache, int32_t kv_heads, int32_t hd,
                        int32_t max_ctx, int32_t pos0, int32_t n, ifb_stream_t st) {
    IFB_CHECK_ARG(out && cache, "GetKRows: null tensor");
    return kv_get_rows(kv_dtype, out, cache, kv_heads, hd, max_ctx, pos0, n, (cudaStream_t)st);
}
int64_t ifb_attn_workspace_bytes(int32_t heads, int32_t hd) { return attn_workspace_bytes(heads, hd); }
int32_t ifb_attn_decode(int32_t kv_dtype, void *out, const void *q, const void *kc, const void *vc,
                        int32_t heads, int32_t kv_heads, int32_t hd, int32_t max_ctx, int32_t ctx_len,
                        float kq_scale, void *ws, ifb_stream_t st) {
    IFB_CHECK_ARG(out && q && kc && vc && ws, "attention: null tensor");
    AttnDecode a;
    a.kv_dtype = kv_dtype; a.heads = heads; a.kv_heads = kv_heads; a.head_dim = hd; a.max_ctx = max_ctx;
    a.ctx_len = ctx_len; a.kq_scale = kq_scale; a.out = out; a.q = q; a.kcache = kc; a.vcache = vc; a.workspace = ws;
    return attn_decode(a, (cudaStream_t)st);
}

int32_t ifb_attn_prefill(int32_t kv_dtype, void *out, const void *q, const void *kc, const void *vc, int32_t tokens, int32_t pos0,
                         int32_t heads, int32_t kv_heads, int32_t hd, int32_t max_ctx, ifb_stream_t st) {
    IFB_CHECK_ARG(out && q && kc && vc, "prefill attention: null tensor");
    return attn_prefill(kv_dtype, out, q, kc, vc, tokens, pos0, heads, kv_heads, hd, max_ctx, (cudaStream_t)st);
}

int32_t ifb_attn_decode_append(int32_t kv_dtype, void *out, const void *q, const void *k_new, const void *v_new,
                               void *kc, void *vc, int32_t heads, int32_t kv_heads, int32_t hd, int32_t max_ctx,
                               int32_t ctx_len, float kq_scale, int32_t rope_order, int32_t rope_dims,
                               float rope_theta, void *ws, ifb_stream_t st) {
    IFB_CHECK_ARG(out && q && k_new && v_new && kc && vc && ws, "attention: null tensor");
    AttnDecode a;
    a.kv_dtype = kv_dtype; a.heads = heads; a.kv_heads = kv_heads; a.head_dim = hd; a.max_ctx = max_ctx;
    a.ctx_len = ctx_len; a.kq_scale = kq_scale; a.out = out; a.q = q; a.kcache = kc; a.vcache = vc; a.workspace = ws;
    a.k_new = k_new; a.v_new = v_new;
    a.rope = rope_order >= 0 ? 1 : 0; a.rope_order = rope_order >= 0 ? rope_order : 0; a.rope_dims = rope_dims;
    a.rope_theta = rope_theta;
    return attn_decode(a, (cudaStream_t)st);
}

int32_t ifb_moe_topk(int32_t *ids, float *w, int32_t *cnt, const void *scores, int32_t rows, int32_t experts,
                     int32_t top_k, int32_t norm, ifb_stream_t st) {
    IFB_CHECK_ARG(ids && w && cnt && scores, "BuildRowsForMoE: null tensor");
    return moe_topk(ids, w, cnt, scores, rows, experts, top_k, norm, (cudaStream_t)st);
}

int32_t ifb_moe_route(int32_t *ids, void *w, const void *logits, int32_t experts, int32_t top_k, int32_t norm,
                      ifb_stream_t st) {
    IFB_CHECK_ARG(ids && w && logits, "moe_route: null tensor");
    return moe_route(ids, w, logits, experts, top_k, norm, (cudaStream_t)st);
}
int32_t ifb_moe_combine(void *x_out, const void *x_res, const void *y, const int32_t *ids, const void *w,
                        int32_t top_k, int64_t d, ifb_stream_t st) {
    IFB_CHECK_ARG(x_out && x_res && y && ids && w, "moe_combine: null tensor");
    return moe_combine(x_out, x_res, y, ids, w, top_k, d, (cudaStream_t)st);
}

}  // extern "C"
