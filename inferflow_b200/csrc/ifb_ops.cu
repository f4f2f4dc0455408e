// Small per-token operators of the decode path (latency-bound; most are fused into the GEMV /
// attention kernels by the model runtime and exist stand-alone for the operator-level API).
//
// Reference: TensorOpr::{RmsNorm, PositionEmbedding, Activation, Mul, Add} (src/tensor/tensor_opr.cu:
// 557,693-760,...; kernels src/kernels/unary_tensor_opr.h:217,553,662,700), SampleTokens greedy
// pick (src/transformer/inference_engine.cc:1986), HostTensorOpr::BuildRowsForMoE
// (src/tensor/host_tensor_opr.cc:190-244).
#include "ifb_formats.cuh"
#include "ifb_kernels.h"

namespace ifb {

// ---------------------------------------------------------------- RMSNorm
// y = x * 1/sqrt(mean(x^2)+eps) * (multi_base + w), one CTA per row
__global__ void __launch_bounds__(256) rmsnorm_kernel(__half *y, const __half *x, const __half *w,
                                                      int64_t cols, float eps, float base) {
    __shared__ float red[8];
    const __half *xr = x + (int64_t)blockIdx.x * cols;
    __half *yr = y + (int64_t)blockIdx.x * cols;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < cols; i += 256) {
        float v = __half2float(xr[i]);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i];
    float scale = 1.0f / sqrtf(t / (float)cols + eps);
    for (int64_t i = threadIdx.x; i < cols; i += 256) {
        float v = __half2float(xr[i]) * scale;
        if (w != nullptr) v = v * (base + __half2float(w[i]));
        yr[i] = __float2half_rn(v);
    }
}

int rmsnorm(void *y, const void *x, const void *w, int64_t rows, int64_t cols, float eps,
            float base, cudaStream_t st) {
    IFB_CHECK_ARG(rows >= 0 && cols > 0, "RmsNorm: bad shape");
    if (rows == 0) return IFB_OK;
    rmsnorm_kernel<<<(unsigned)rows, 256, 0, st>>>((__half *)y, (const __half *)x, (const __half *)w,
                                                   cols, eps, base);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- RoPE
// angle_i = pos * theta^(-2i/dims) computed as in the reference: theta_scale = powf(theta,-2/dims),
// angle = pos * powf(theta_scale, i) (unary_tensor_opr.h:666-679,706-719)
__device__ __forceinline__ void rope_pair(float &x0, float &x1, int pos, int i, float theta_scale) {
    float ang = (float)pos;
    if (i > 0) ang *= powf(theta_scale, (float)i);
    float c = cosf(ang), s = sinf(ang);
    float a = x0 * c - x1 * s, b = x0 * s + x1 * c;
    x0 = a;
    x1 = b;
}

__global__ void rope_kernel(__half *x, int tokens, int heads, int head_dim, int rope_dims, int pos0,
                            float theta, int order) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int pairs = rope_dims / 2;
    int total = tokens * heads * pairs;
    if (idx >= total) return;
    int i = idx % pairs, h = (idx / pairs) % heads, t = idx / (pairs * heads);
    float theta_scale = powf(theta, -2.0f / (float)rope_dims);
    __half *v = x + ((int64_t)t * heads + h) * head_dim;
    int i0 = order == 2 ? i : 2 * i, i1 = order == 2 ? i + rope_dims / 2 : 2 * i + 1;
    float x0 = __half2float(v[i0]), x1 = __half2float(v[i1]);
    rope_pair(x0, x1, pos0 + t, i, theta_scale);
    v[i0] = __float2half_rn(x0);
    v[i1] = __float2half_rn(x1);
}

int rope(void *x, int tokens, int heads, int head_dim, int rope_dims, int pos0, float theta,
         int order, cudaStream_t st) {
    if (rope_dims <= 0) rope_dims = head_dim;
    IFB_CHECK_ARG(rope_dims % 2 == 0 && rope_dims <= head_dim, "RoPE: bad rope_dims %d", rope_dims);
    IFB_CHECK_ARG(order == 0 || order == 2, "RoPE: order %d (0 = interleaved, 2 = half-split)", order);
    int total = tokens * heads * (rope_dims / 2);
    if (total <= 0) return IFB_OK;
    rope_kernel<<<(total + 127) / 128, 128, 0, st>>>((__half *)x, tokens, heads, head_dim, rope_dims,
                                                     pos0, theta, order);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- elementwise
__global__ void silu_mul_kernel(__half *o, const __half *a, const __half *b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = __half2float(a[i]);
    float fx = round_h(x / (1.0f + expf(-x)));  // SiLU output tensor is half
    o[i] = __float2half_rn(fx * __half2float(b[i]));
}
__global__ void add_kernel(__half *c, const __half *a, const __half *b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    c[i] = __float2half_rn(__half2float(a[i]) + __half2float(b[i]));
}
int silu_mul(void *o, const void *a, const void *b, int64_t n, cudaStream_t st) {
    if (n <= 0) return IFB_OK;
    silu_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((__half *)o, (const __half *)a,
                                                                 (const __half *)b, n);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}
int add(void *c, const void *a, const void *b, int64_t n, cudaStream_t st) {
    if (n <= 0) return IFB_OK;
    add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((__half *)c, (const __half *)a,
                                                            (const __half *)b, n);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- argmax (first max wins)
// ll != nullptr: the logits arrive as flag-in-data words {half2, flag} from the vocabulary-sharded lm_head of every
// rank (ifb_gemv.cu, tensor-parallel push); wait for each word, store the compact fp16 logits, take the argmax.
__global__ void __launch_bounds__(1024) argmax_kernel(int32_t *out, const __half *x, int64_t n, __half *compact,
                                                      const uint2 *ll, const int32_t *seq, uint32_t idx) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -INFINITY;
    int bi = 0x7fffffff;   // "none yet": mapped to token 0 at the end (all-NaN logits must not index past the vocabulary)
    if (ll != nullptr) {
        const uint32_t flag = ((uint32_t)*reinterpret_cast<const volatile int32_t *>(seq) << 8) | idx;
        for (int64_t w = threadIdx.x; w < n / 2; w += 1024) {
            uint2 v = ld_relaxed_sys_v2(ll + w);
            PeerDeadline dl;   // bounded: see ifb_common.cuh
            while (v.y != flag) { v = ld_relaxed_sys_v2(ll + w); dl.tick(); }
            reinterpret_cast<uint32_t *>(compact)[w] = v.x;
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
            if (f.x > best) { best = f.x; bi = (int)(2 * w); }       // ascending index per thread: first max wins
            if (f.y > best) { best = f.y; bi = (int)(2 * w + 1); }
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += 1024) {
            float v = __half2float(x[i]);
            if (v > best || (v == best && (int)i < bi)) { best = v; bi = (int)i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = sv[threadIdx.x];
        bi = si[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) *out = bi == 0x7fffffff ? 0 : bi;
    }
}
int argmax(int32_t *out, const void *logits, int64_t n, cudaStream_t st) {
    IFB_CHECK_ARG(n > 0, "argmax: empty input");
    argmax_kernel<<<1, 1024, 0, st>>>(out, (const __half *)logits, n, nullptr, nullptr, nullptr, 0);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}
int argmax_tp(int32_t *out, void *logits_f16, const void *ll_words, int64_t n, const int32_t *seq, uint32_t idx,
              cudaStream_t st) {
    IFB_CHECK_ARG(n > 0 && n % 2 == 0 && ll_words != nullptr && seq != nullptr, "argmax_tp: bad arguments");
    argmax_kernel<<<1, 1024, 0, st>>>(out, nullptr, n, (__half *)logits_f16, (const uint2 *)ll_words, seq, idx);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- embedding row
// GetEmbdTensor (inference_engine.cc:1324-1350): copy / dequantize one row of the table to fp16
__global__ void embed_row_kernel(__half *dst, const uint8_t *table, int dt, int64_t cols,
                                 int token_id, const int32_t *token_dev, int rows) {
    int tok = token_dev != nullptr ? *token_dev : token_id;
    if (rows > 0) tok = max(0, min(tok, rows - 1));   // a token id from device memory is not validated by the host
    if (dt == DT_F16) {
        const __half *row = reinterpret_cast<const __half *>(table) + (int64_t)tok * cols;
        for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += (int64_t)gridDim.x * blockDim.x)
            dst[i] = row[i];
    } else if (dt == DT_F32) {
        const float *row = reinterpret_cast<const float *>(table) + (int64_t)tok * cols;
        for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += (int64_t)gridDim.x * blockDim.x)
            dst[i] = __float2half_rn(row[i]);
    } else {
        int cap = block_capacity(dt), bb = block_bytes(dt);
        const uint8_t *row = table + (int64_t)tok * row_bytes(dt, cols);
        for (int64_t b = blockIdx.x * blockDim.x + threadIdx.x; b < cols / cap; b += (int64_t)gridDim.x * blockDim.x) {
            __half *o = dst + b * cap;
            dequant_block(dt, row + b * bb, [&](int i, float v) { o[i] = __float2half_rn(v); });
        }
    }
}
int embed_row(void *dst, const void *table, int dt, int64_t cols, int token_id,
              const int32_t *token_dev, cudaStream_t st, int rows) {
    embed_row_kernel<<<4, 256, 0, st>>>((__half *)dst, (const uint8_t *)table, dt, cols, token_id, token_dev, rows);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- MoE router top-k
// HostTensorOpr::BuildRowsForMoE (host_tensor_opr.cc:190-244): insertion scan in expert order,
// first wins on ties, scores < 1e-5 dropped, optional renormalisation. One thread per row
// (rows <= 16 queries, experts <= 64: latency only). ids are bit-exact with the host rule.
__global__ void moe_topk_kernel(int32_t *ids, float *wts, int32_t *cnt, const __half *scores,
                                int rows, int experts, int top_k, int norm) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const __half *s = scores + (int64_t)r * experts;
    int32_t id[8];
    float wt[8];
    int n = 0;
    for (int e = 0; e < experts; e++) {
        float sc = __half2float(s[e]);
        if ((n >= top_k && wt[n - 1] >= sc) || sc < 0.00001f) continue;
        if (n < top_k) { id[n] = -1; wt[n] = 0.f; n++; }
        for (int j = n - 1; j >= 0; j--) {
            if (wt[j] >= sc) break;
            if (j + 1 < n) { id[j + 1] = id[j]; wt[j + 1] = wt[j]; }
            id[j] = e;
            wt[j] = sc;
        }
    }
    if (norm) {
        float sum = 0.f;
        for (int j = 0; j < n; j++) sum += wt[j];
        for (int j = 0; j < n; j++) wt[j] = __fdiv_rn(wt[j], sum);
    }
    cnt[r] = n;
    for (int j = 0; j < top_k; j++) {
        ids[(int64_t)r * top_k + j] = j < n ? id[j] : -1;
        wts[(int64_t)r * top_k + j] = j < n ? wt[j] : 0.f;
    }
}
int moe_topk(int32_t *ids, float *weights, int32_t *counts, const void *scores, int rows,
             int experts, int top_k, int norm, cudaStream_t st) {
    IFB_CHECK_ARG(top_k >= 1 && experts >= 1, "moe_topk: bad sizes");
    IFB_CHECK_ARG(top_k <= 8, "moe_topk: top_k %d > RowItemForMoe::MAX_SIZE (8)", top_k);
    if (rows <= 0) return IFB_OK;
    moe_topk_kernel<<<(rows + 31) / 32, 32, 0, st>>>(ids, weights, counts, (const __half *)scores,
                                                     rows, experts, top_k, norm);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------- MoE routing of one token, all on the device
// ProcessGpuLayer_Moe (inference_worker.cc:1965-2010): SoftMax over ALL experts (Tensor_SoftMax_Alg2_Kernel,
// unary_tensor_opr.h:481: fp32 max / exp / sum over a warp, the exponentials stored as half, then multiplied by
// half(1/sum) in half), D2H, BuildRowsForMoE on the host, experts visited in ascending id. Here: one warp, no D2H.
__global__ void moe_route_kernel(int32_t *ids, __half *wts, const __half *logits, int experts, int top_k, int norm) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x;
    __shared__ float s_p[32];
    float v = lane < experts ? __half2float(logits[lane]) : -INFINITY;
    float mx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float ex = lane < experts ? expf(v - mx) : 0.f;
    float sum = ex;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const __half ph = __hmul(__float2half_rn(ex), __float2half_rn(1.0f / sum));
    s_p[lane] = lane < experts ? __half2float(ph) : 0.f;
    __syncwarp();
    if (lane != 0) return;
    int id[8];
    float wt[8];
    int n = 0;
    for (int e = 0; e < experts; e++) {       // HostTensorOpr::BuildRowsForMoE (host_tensor_opr.cc:190-244)
        const float sc = s_p[e];
        if ((n >= top_k && wt[n - 1] >= sc) || sc < 0.00001f) continue;
        if (n < top_k) { id[n] = -1; wt[n] = 0.f; n++; }
        for (int j = n - 1; j >= 0; j--) {
            if (wt[j] >= sc) break;
            if (j + 1 < n) { id[j + 1] = id[j]; wt[j + 1] = wt[j]; }
            id[j] = e;
            wt[j] = sc;
        }
    }
    if (norm) {
        float t = 0.f;
        for (int j = 0; j < n; j++) t += wt[j];
        for (int j = 0; j < n; j++) wt[j] = __fdiv_rn(wt[j], t);
    }
    for (int i = 1; i < n; i++) {             // ascending expert id: the order the reference runs the experts in
        const int ci = id[i];
        const float cw = wt[i];
        int j = i - 1;
        while (j >= 0 && id[j] > ci) { id[j + 1] = id[j]; wt[j + 1] = wt[j]; j--; }
        id[j + 1] = ci;
        wt[j + 1] = cw;
    }
    for (int j = 0; j < top_k; j++) {
        ids[j] = j < n ? id[j] : -1;
        wts[j] = __float2half_rn(j < n ? wt[j] : 0.f);   // aux_buffer_ holds the weights as half
    }
}
int moe_route(int32_t *ids, void *weights_f16, const void *logits_f16, int experts, int top_k, int norm, cudaStream_t st) {
    IFB_CHECK_ARG(experts >= 1 && experts <= 32, "moe_route: %d experts (1..32)", experts);
    IFB_CHECK_ARG(top_k >= 1 && top_k <= 8 && top_k <= experts, "moe_route: top_k %d", top_k);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    IFB_CUDA(cudaLaunchKernelEx(&cfg, moe_route_kernel, ids, (__half *)weights_f16, (const __half *)logits_f16, experts, top_k, norm));
    return IFB_OK;
}

__global__ void moe_combine_kernel(__half *x_out, const __half *x_res, const __half *y, const int32_t *ids,
                                   const __half *wts, int top_k, int64_t d) {
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        __half out = __float2half_rn(0.f);                       // out_tensor->AssignZero()
        for (int k = 0; k < top_k; k++)
            if (ids[k] >= 0) out = __hfma(y[(int64_t)k * d + i], wts[k], out);   // AddByRowIdx_Kernel: B = fma(A, w, B) in half
        x_out[i] = __hadd(x_res[i], out);                        // the layer's residual Add
    }
}
// tensor-parallel form: the experts' w2 are row-parallel, so y holds this rank's PARTIAL sums; the weighted combination
// is linear, each rank folds its partials and pushes the result into every rank's exchange region as flag-in-data words
// {half2, flag}; the next GEMV's prologue adds the ranks' vectors in rank order and the residual last (tp_gather8).
__global__ void moe_combine_push_kernel(const __half *y, const int32_t *ids, const __half *wts, int top_k, int64_t d,
                                        TpExchange tp) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t flag = ((uint32_t)*reinterpret_cast<const volatile int32_t *>(tp.seq) << 8) | tp.push_idx;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < d / 2; w += (int64_t)gridDim.x * blockDim.x) {
        __half2 out = __floats2half2_rn(0.f, 0.f);
        for (int k = 0; k < top_k; k++)
            if (ids[k] >= 0) out = __hfma2(reinterpret_cast<const __half2 *>(y + (int64_t)k * d)[w], __half2half2(wts[k]), out);
        const uint2 word = make_uint2(*reinterpret_cast<const uint32_t *>(&out), flag);
        for (int p = 0; p < tp.size; p++)
            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(reinterpret_cast<uint8_t *>(tp.peer[p]) + tp.dst_off + w * 8),
                         "r"(word.x), "r"(word.y) : "memory");
    }
}
int moe_combine(void *x_out, const void *x_res, const void *y, const int32_t *ids, const void *weights_f16, int top_k,
                int64_t d, cudaStream_t st, const TpExchange *tp) {
    IFB_CHECK_ARG(top_k >= 1 && top_k <= 8 && d > 0, "moe_combine: bad sizes");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((d + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (tp != nullptr && tp->push) {
        IFB_CHECK_ARG(d % 2 == 0, "moe_combine: odd vector length");
        IFB_CUDA(cudaLaunchKernelEx(&cfg, moe_combine_push_kernel, (const __half *)y, ids, (const __half *)weights_f16, top_k, d, *tp));
        return IFB_OK;
    }
    IFB_CUDA(cudaLaunchKernelEx(&cfg, moe_combine_kernel, (__half *)x_out, (const __half *)x_res, (const __half *)y, ids,
                                (const __half *)weights_f16, top_k, d));
    return IFB_OK;
}

}  // namespace ifb
