// KV cache and single-token grouped-query attention, sm_100a.
//
// Replaces GpuLayerKVCache::{SetKRows,SetVRows,GetKRows,GetVRows} (reference
// src/transformer/kv_cache.cc:104-249) and the attention sub-graph of
// GpuInferenceWorker::ProcessGpuLayer_Attention (src/transformer/inference_worker.cc:1130-1320,
// 1587-1724): RoPE on q/k, append, [dequantize whole K prefix -> RepeatKV -> Gemm_Alg2(alpha) ->
// SoftMax(causal, kq_scale) -> dequantize whole V -> Transpose -> RepeatKV -> Gemm_Alg2 -> Assign].
//
// Here: the cache is read IN PLACE exactly once per token (flash-decoding): split over the
// context, K/V tiles staged global->shared by 1-D TMA bulk copies (cp.async.bulk) behind
// mbarriers, online softmax in fp32 with warp-shuffle reductions, GQA by sharing the staged
// tile across the q-heads of the group (no RepeatKV), split partials merged by the last CTA.
//
// Cache layout (ours; the VALUES -- int8 code and fp16 scale per 32 consecutive dims of a
// row -- are the reference's Q8_B32T2 rows, kernels/tensor_quant.h:45-82):
//   F16      : half [kv_heads][max_ctx][head_dim]
//   Q8_B32T2 : int8 [kv_heads][max_ctx][head_dim] ++ half [kv_heads][max_ctx][head_dim/32]
// head-major so that a (head, position-range) tile is ONE contiguous byte range.
#include "ifb_formats.cuh"
#include "ifb_kernels.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace ifb {

constexpr int ATT_CHUNK = 64;    // the cache's position padding (layout constant)
constexpr int ATT_TILE = 32;     // positions per staged tile: one per lane in the score phase
constexpr int ATT_THREADS = 256;

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int64_t kv_cache_bytes(int kv_dtype, int kv_heads, int head_dim, int max_ctx) {
    int64_t rows = (int64_t)kv_heads * round_up(max_ctx, ATT_CHUNK);
    if (kv_dtype == DT_F16) return rows * head_dim * 2;
    if (kv_dtype == DT_Q8_B32T2) return rows * head_dim + rows * (head_dim / 32) * 2;
    return -1;
}

// ---------------------------------------------------------------------------------------
// append (+ fused RoPE on q and k)
// ---------------------------------------------------------------------------------------
struct KvArgs {
    int kv_dtype, kv_heads, head_dim, max_ctx_pad, pos0, n_tokens;
    uint8_t *kcache, *vcache;
    const __half *k, *v;
    int rope, rope_order, rope_dims;
    float rope_theta;
    __half *q;
    int q_heads;
    const int32_t *pos_dev;
};

__device__ __forceinline__ void rope_rot(float &x0, float &x1, int pos, int i, float theta_scale) {
    float ang = (float)pos;
    if (i > 0) ang *= powf(theta_scale, (float)i);
    float c = cosf(ang), s = sinf(ang);
    float a = x0 * c - x1 * s, b = x0 * s + x1 * c;
    x0 = a;
    x1 = b;
}

// quantize / store one 32-dim block of a K or V head row: lane = element
__device__ __forceinline__ void kv_store_block(int kv_dtype, uint8_t *cache, float v, int kv_heads,
                                               int hd, int max_ctx_pad, int h, int pos, int blk, int lane) {
    const int64_t base = (int64_t)h * max_ctx_pad + pos;
    const int dh = blk * 32 + lane;
    if (kv_dtype == DT_F16) {
        reinterpret_cast<__half *>(cache)[base * hd + dh] = __float2half_rn(v);
    } else {
        // Tensor_QuantizeQ8_B32T2_Alg2_Kernel (tensor_quant.h:45-82)
        int8_t *codes = reinterpret_cast<int8_t *>(cache);
        __half *scales = reinterpret_cast<__half *>(cache + (int64_t)kv_heads * max_ctx_pad * hd);
        const float m = warp_max(fabsf(v));
        const float scale = __fdiv_rn(m, 127.0f);
        int q = scale <= 0.000001f ? 0 : (int)roundf(__fdiv_rn(v, scale));
        q = min(max(q, -128), 127);
        codes[base * hd + dh] = (int8_t)q;
        if (lane == 0) scales[base * (hd / 32) + blk] = __float2half_rn(scale);
    }
}

// grid (kv_heads, n_tokens), 128 threads: CTA = one kv head of one token and the q heads of its group
__global__ void __launch_bounds__(128) kv_append_kernel(const KvArgs a) {
    __shared__ float s_cos[128], s_sin[128];
    __shared__ float s_k[256];
    pdl_launch_dependents();
    const int h = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the position is stable before this token's first kernel starts: the rotation table does not
    // depend on the QKV GEMV and is built while that kernel is still running
    const int pos = (a.pos_dev != nullptr ? *a.pos_dev : a.pos0) + t;
    const int hd = a.head_dim, kvd = a.kv_heads * hd;
    const __half *k = a.k + (int64_t)t * kvd + (int64_t)h * hd;
    const __half *v = a.v + (int64_t)t * kvd + (int64_t)h * hd;
    const int rd = a.rope_dims > 0 ? a.rope_dims : hd;
    const int pairs = rd / 2;
    if (a.rope) {
        // angle_i = pos * theta_scale^i, theta_scale = theta^(-2/dims) (unary_tensor_opr.h:666-679,706-719)
        const float theta_scale = powf(a.rope_theta, -2.0f / (float)rd);
        for (int i = tid; i < pairs; i += 128) {
            float ang = (float)pos;
            if (i > 0) ang *= powf(theta_scale, (float)i);
            s_cos[i] = cosf(ang);
            s_sin[i] = sinf(ang);
        }
    }
    pdl_wait();

    for (int i = tid; i < hd; i += 128) s_k[i] = __half2float(k[i]);
    if (a.rope) {
        __syncthreads();
        for (int i = tid; i < pairs; i += 128) {
            const int i0 = a.rope_order == 2 ? i : 2 * i, i1 = a.rope_order == 2 ? i + rd / 2 : 2 * i + 1;
            const float x0 = s_k[i0], x1 = s_k[i1], c = s_cos[i], sn = s_sin[i];
            s_k[i0] = round_h(x0 * c - x1 * sn);   // PositionEmbedding writes half
            s_k[i1] = round_h(x0 * sn + x1 * c);
        }
        if (a.q != nullptr) {
            const int G = a.q_heads / a.kv_heads;
            __half *q = a.q + (int64_t)t * a.q_heads * hd + (int64_t)h * G * hd;
            for (int idx = tid; idx < G * pairs; idx += 128) {
                const int i = idx % pairs, g = idx / pairs;
                const int i0 = a.rope_order == 2 ? i : 2 * i, i1 = a.rope_order == 2 ? i + rd / 2 : 2 * i + 1;
                const float x0 = __half2float(q[g * hd + i0]), x1 = __half2float(q[g * hd + i1]);
                const float c = s_cos[i], sn = s_sin[i];
                q[g * hd + i0] = __float2half_rn(x0 * c - x1 * sn);
                q[g * hd + i1] = __float2half_rn(x0 * sn + x1 * c);
            }
        }
    }
    __syncthreads();
    const int nb = hd / 32;
    for (int b = warp; b < 2 * nb; b += 4) {
        if (b < nb) kv_store_block(a.kv_dtype, a.kcache, s_k[b * 32 + lane], a.kv_heads, hd, a.max_ctx_pad, h, pos, b, lane);
        else kv_store_block(a.kv_dtype, a.vcache, __half2float(v[(b - nb) * 32 + lane]), a.kv_heads, hd, a.max_ctx_pad, h, pos, b - nb, lane);
    }
}

int kv_append(const KvAppend &p, cudaStream_t st) {
    IFB_CHECK_ARG(p.kv_dtype == DT_F16 || p.kv_dtype == DT_Q8_B32T2,
                  "kv cache type %d (F16 or Q8_B32T2 only, inference_engine.cc:1702-1704)", p.kv_dtype);
    IFB_CHECK_ARG(p.head_dim == 32 || p.head_dim == 64 || p.head_dim == 128, "kv: head_dim %d not supported (32/64/128)", p.head_dim);
    IFB_CHECK_ARG(p.pos_dev != nullptr || (p.pos0 >= 0 && p.pos0 + p.n_tokens <= p.max_ctx),
                  "kv: rows %d..%d outside max_context_len %d", p.pos0, p.pos0 + p.n_tokens, p.max_ctx);
    if (p.n_tokens <= 0) return IFB_OK;
    KvArgs a;
    a.kv_dtype = p.kv_dtype; a.kv_heads = p.kv_heads; a.head_dim = p.head_dim;
    a.max_ctx_pad = round_up(p.max_ctx, ATT_CHUNK); a.pos0 = p.pos0; a.n_tokens = p.n_tokens;
    a.kcache = (uint8_t *)p.kcache; a.vcache = (uint8_t *)p.vcache;
    a.k = (const __half *)p.k; a.v = (const __half *)p.v;
    a.rope = p.rope; a.rope_order = p.rope_order; a.rope_dims = p.rope_dims; a.rope_theta = p.rope_theta;
    a.q = (__half *)p.q; a.q_heads = p.q_heads; a.pos_dev = p.pos_dev;
    IFB_CHECK_ARG(p.q == nullptr || (p.q_heads > 0 && p.q_heads % p.kv_heads == 0), "kv: q_heads %d / kv_heads %d", p.q_heads, p.kv_heads);
    IFB_CHECK_ARG(p.n_tokens <= 65535, "kv: too many tokens in one append");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.kv_heads, p.n_tokens); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = p.pdl ? 1 : 0;
    IFB_CUDA(cudaLaunchKernelEx(&cfg, kv_append_kernel, a));
    return IFB_OK;
}

// rows pos0..pos0+n-1 as fp16 [n][kv_heads*head_dim] (GetKRows: dequantized view)
__global__ void kv_get_rows_kernel(__half *out, const uint8_t *cache, int kv_dtype, int kv_heads,
                                   int hd, int max_ctx_pad, int pos0, int n) {
    int64_t total = (int64_t)n * kv_heads * hd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int d = (int)(i % (kv_heads * hd));
        int t = (int)(i / (kv_heads * hd));
        int h = d / hd, dh = d % hd;
        int64_t base = (int64_t)h * max_ctx_pad + pos0 + t;
        if (kv_dtype == DT_F16) {
            out[i] = reinterpret_cast<const __half *>(cache)[base * hd + dh];
        } else {
            const int8_t *codes = reinterpret_cast<const int8_t *>(cache);
            const __half *scales = reinterpret_cast<const __half *>(cache + (int64_t)kv_heads * max_ctx_pad * hd);
            out[i] = __float2half_rn((float)codes[base * hd + dh] * __half2float(scales[base * (hd / 32) + dh / 32]));
        }
    }
}

int kv_get_rows(int kv_dtype, void *out, const void *cache, int kv_heads, int head_dim, int max_ctx,
                int pos0, int n, cudaStream_t st) {
    IFB_CHECK_ARG(kv_dtype == DT_F16 || kv_dtype == DT_Q8_B32T2, "kv cache type %d", kv_dtype);
    IFB_CHECK_ARG(pos0 >= 0 && n >= 0 && pos0 + n <= max_ctx, "kv: rows outside the cache");
    if (n == 0) return IFB_OK;
    kv_get_rows_kernel<<<256, 256, 0, st>>>((__half *)out, (const uint8_t *)cache, kv_dtype, kv_heads,
                                            head_dim, round_up(max_ctx, ATT_CHUNK), pos0, n);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// ---------------------------------------------------------------------------------------
// decode attention
// ---------------------------------------------------------------------------------------
struct AttnArgs {
    int heads, kv_heads, max_ctx_pad, ctx_len, nsplit;
    float softmax_scale;  // 1/sqrt(head_dim)  (= alpha * kq_scale of the reference)
    __half *out;
    const __half *q;
    const uint8_t *kcache, *vcache;
    float *ws;            // [heads][nsplit][HD+2]
    unsigned int *counters;  // [kv_heads]
    const int32_t *pos_dev;
    unsigned long long *trace;
    int cluster;          // the nsplit CTAs of a kv head form one thread-block cluster: merge through DSMEM
    int part_off;         // byte offset (dynamic shared memory) of the leader's [nsplit][G][HD+2] partials
    // fused append
    const __half *k_new, *v_new;
    int fuse, rope, rope_order, rope_dims;
    float rope_theta;
};

int64_t attn_workspace_bytes(int heads, int head_dim) {
    // worst-case split count = 2*SMs (one per CTA) ; + counters
    int64_t max_split = 4 * 160;
    return (int64_t)heads * max_split * (head_dim + 2) * sizeof(float) + 4096;
}

// signed byte k of w -> float without I2F: flip the sign bit (u = b + 128), drop the byte into the
// mantissa of 2^23 with one PRMT, subtract 2^23 + 128. Exact.
__device__ __forceinline__ float s8f(uint32_t w_biased, int k) {
    const uint32_t bits = __byte_perm(w_biased, 0x4b000000u, 0x7650u + (uint32_t)k);
    return __uint_as_float(bits) - 8388736.0f;
}
__device__ __forceinline__ uint32_t s8bias(uint32_t w) { return w ^ 0x80808080u; }

// Each WARP owns whole tiles of 64 positions (its own shared-memory buffer, its own mbarrier, its
// own online softmax): no block-wide barrier inside the scan. The 4 warps of a CTA merge through
// shared memory, CTAs of one kv head through the workspace (last CTA merges, log-sum-exp).
template <int HD, int G, bool Q8>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const AttnArgs a) {
    constexpr int NB = HD / 32;                       // scale blocks per (pos, head)
    constexpr int KBYTES = Q8 ? ATT_TILE * HD : ATT_TILE * HD * 2;
    constexpr int SBYTES = Q8 ? ATT_TILE * NB * 2 : 0;
    constexpr int TILE = 2 * (KBYTES + SBYTES);       // K codes, K scales, V codes, V scales
    constexpr int LPP = HD / 16;                      // lanes per position in the QK phase
    constexpr int PPW = 32 / LPP;                     // positions per warp iteration
    constexpr int NW = ATT_THREADS / 32;

    extern __shared__ __align__(128) uint8_t smem[];  // [NW][TILE]
    __shared__ uint64_t full_bar[NW];
    __shared__ float s_p[NW][G][ATT_TILE];           // scores / probabilities of the warp's tile
    __shared__ __align__(16) float s_q[G][HD / 16][20];   // 16 dims + 4 pad: conflict-free LDS.128
    // Q8 cache: q as TWO int8 terms per 32-dim block, q ~= s1*qh + s2*ql (s2 <= s1/254: ~15 bits, finer than the
    // fp16 q it is made from), so q.k over int8 K codes is two dp4a per 4 dims instead of 4 converts + 4 FMAs
    __shared__ __align__(16) int8_t s_qh[G][HD], s_ql[G][HD];
    __shared__ float s_qs[G][NB][2];
    __shared__ float s_ml[NW][G][2];
    __shared__ unsigned int s_last;
    __shared__ float s_cos[HD / 2], s_sin[HD / 2];          // fused RoPE table of this position
    __shared__ float s_new[2][HD];                          // the new K (after RoPE) and V rows, fp16-rounded
    __shared__ __align__(16) uint8_t s_newrow[2][HD * 2];   // ... as stored in the cache (codes or halves)
    __shared__ __half s_newsc[2][NB];                       // ... and their block scales (Q8)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kvh = blockIdx.y, split = blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < NW; i++) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    if (tid == 0) trace_mark(a.trace, 0);
    // The position was written before this token's first kernel started (it is stable here), and
    // every cache row except the NEW one (position ctx-1, written by the kv-append kernel just
    // ahead of us) is older than the previous token: those tiles are fetched BEFORE the grid
    // dependency resolves, i.e. while the QKV GEMV is still running.
    const int ctx = a.pos_dev != nullptr ? (*a.pos_dev + 1) : a.ctx_len;
    const int nchunks_total = (ctx + ATT_TILE - 1) / ATT_TILE;
    const int cps = (nchunks_total + a.nsplit - 1) / a.nsplit;   // chunks per CTA
    const int c0 = split * cps, c1 = min(c0 + cps, nchunks_total);
    const int new_chunk = (ctx - 1) / ATT_TILE;                 // the tile holding the new row

    const int64_t head_rows = (int64_t)kvh * a.max_ctx_pad;
    const uint8_t *kc = a.kcache + (Q8 ? head_rows * HD : head_rows * HD * 2);
    const uint8_t *vc = a.vcache + (Q8 ? head_rows * HD : head_rows * HD * 2);
    const uint8_t *ks = a.kcache + (int64_t)a.kv_heads * a.max_ctx_pad * HD + head_rows * NB * 2;
    const uint8_t *vs = a.vcache + (int64_t)a.kv_heads * a.max_ctx_pad * HD + head_rows * NB * 2;
    uint8_t *tile = smem + (size_t)warp * TILE;

    auto issue = [&](int c) {  // tile c of the sequence -> this warp's buffer
        const int64_t p0 = (int64_t)c * ATT_TILE;
        mbar_arrive_expect_tx(&full_bar[warp], TILE);
        if (Q8) {
            bulk_g2s(tile, kc + p0 * HD, KBYTES, &full_bar[warp]);
            bulk_g2s(tile + KBYTES, ks + p0 * NB * 2, SBYTES, &full_bar[warp]);
            bulk_g2s(tile + KBYTES + SBYTES, vc + p0 * HD, KBYTES, &full_bar[warp]);
            bulk_g2s(tile + 2 * KBYTES + SBYTES, vs + p0 * NB * 2, SBYTES, &full_bar[warp]);
        } else {
            bulk_g2s(tile, kc + p0 * HD * 2, KBYTES, &full_bar[warp]);
            bulk_g2s(tile + KBYTES, vc + p0 * HD * 2, KBYTES, &full_bar[warp]);
        }
    };
    int c = c0 + warp;                 // this warp's tiles: c0+warp, c0+warp+NW, ...
    bool issued = false;
    // prefetch paced to one tile in flight per CTA (warp w goes after warp w-1's tile has landed):
    // a deep backlog on the SM's memory port would only delay the running kernels' small loads
    // (fused append: the tile with the new row can be fetched early too -- its stale row is patched in shared
    //  memory from s_newrow once the tile has landed)
    if (lane == 0 && c < c1 && (c != new_chunk || a.fuse)) {
        issue(c);
        issued = true;
    }
    const int rd = a.rope_dims > 0 ? a.rope_dims : HD;
    const int pairs = rd / 2;
    if (a.fuse && a.rope) {
        // angle_i = pos * theta_scale^i, theta_scale = theta^(-2/dims) (unary_tensor_opr.h:666-679,706-719);
        // the position is known before the QKV GEMV has finished
        const float theta_scale = powf(a.rope_theta, -2.0f / (float)rd);
        for (int i = tid; i < pairs; i += ATT_THREADS) {
            float ang = (float)(ctx - 1);
            if (i > 0) ang *= powf(theta_scale, (float)i);
            s_cos[i] = cosf(ang);
            s_sin[i] = sinf(ang);
        }
    }

    pdl_wait();  // q and the new K/V row come from the previous kernels
    if (tid == 0) trace_mark(a.trace, 1);
    if (lane == 0 && c < c1 && !issued) issue(c);
    if (a.fuse && a.rope) __syncthreads();   // rotation table complete

    // stage q (scaled) for the G heads of this group
    if (a.fuse && a.rope) {
        // PositionEmbedding writes q back as half (tensor_opr.cu:693-760): rotate, round, then scale
        for (int i = tid; i < G * HD; i += ATT_THREADS) {
            const int g = i / HD, d = i % HD;
            const __half *qh = a.q + (int64_t)(kvh * G + g) * HD;
            float v = __half2float(qh[d]);
            int pi = -1, other = 0;
            bool first = false;
            if (a.rope_order == 2) {
                if (d < rd / 2) { pi = d; other = d + rd / 2; first = true; }
                else if (d < rd) { pi = d - rd / 2; other = d - rd / 2; }
            } else if (d < rd) {
                pi = d >> 1; other = d ^ 1; first = (d & 1) == 0;
            }
            if (pi >= 0) {
                const float o = __half2float(qh[other]), cs = s_cos[pi], sn = s_sin[pi];
                v = first ? round_h(v * cs - o * sn) : round_h(o * sn + v * cs);
            }
            s_q[g][d / 16][d % 16] = v * a.softmax_scale;
        }
    } else {
        for (int i = tid; i < G * HD; i += ATT_THREADS) {
            int g = i / HD, d = i % HD;
            s_q[g][d / 16][d % 16] = __half2float(a.q[(int64_t)(kvh * G + g) * HD + d]) * a.softmax_scale;
        }
    }
    // the CTA whose range holds the new position builds the new K/V row: RoPE(k), quantize, store to the cache
    // (SetKRows / SetVRows, kv_cache.cc:104-183) and keep it in shared memory for its own tile
    const bool owns_new = a.fuse && new_chunk >= c0 && new_chunk < c1;
    if (owns_new) {
        const __half *kn = a.k_new + (int64_t)kvh * HD, *vn = a.v_new + (int64_t)kvh * HD;
        for (int d = tid; d < HD; d += ATT_THREADS) {
            float v = __half2float(kn[d]);
            if (a.rope) {
                int pi = -1, other = 0;
                bool first = false;
                if (a.rope_order == 2) {
                    if (d < rd / 2) { pi = d; other = d + rd / 2; first = true; }
                    else if (d < rd) { pi = d - rd / 2; other = d - rd / 2; }
                } else if (d < rd) {
                    pi = d >> 1; other = d ^ 1; first = (d & 1) == 0;
                }
                if (pi >= 0) {
                    const float o = __half2float(kn[other]), cs = s_cos[pi], sn = s_sin[pi];
                    v = first ? round_h(v * cs - o * sn) : round_h(o * sn + v * cs);
                }
            }
            s_new[0][d] = v;
            s_new[1][d] = __half2float(vn[d]);
        }
        __syncthreads();
        const int pos = ctx - 1;
        const int64_t base = head_rows + pos;
        for (int b = warp; b < 2 * NB; b += NW) {
            const int which = b / NB, blk = b % NB;
            const float v = s_new[which][blk * 32 + lane];
            uint8_t *cache = const_cast<uint8_t *>(which == 0 ? a.kcache : a.vcache);
            if (Q8) {
                // Tensor_QuantizeQ8_B32T2_Alg2_Kernel (tensor_quant.h:45-82)
                const float m = warp_max(fabsf(v));
                const float scale = __fdiv_rn(m, 127.0f);
                int q = scale <= 0.000001f ? 0 : (int)roundf(__fdiv_rn(v, scale));
                q = min(max(q, -128), 127);
                s_newrow[which][blk * 32 + lane] = (uint8_t)(int8_t)q;
                reinterpret_cast<int8_t *>(cache)[base * HD + blk * 32 + lane] = (int8_t)q;
                if (lane == 0) {
                    const __half hs = __float2half_rn(scale);
                    s_newsc[which][blk] = hs;
                    reinterpret_cast<__half *>(cache + (int64_t)a.kv_heads * a.max_ctx_pad * HD)[base * NB + blk] = hs;
                }
            } else {
                const __half hv = __float2half_rn(v);
                reinterpret_cast<__half *>(s_newrow[which])[blk * 32 + lane] = hv;
                reinterpret_cast<__half *>(cache)[base * HD + blk * 32 + lane] = hv;
            }
        }
    }
    __syncthreads();
    if (Q8) {
        for (int b = warp; b < G * NB; b += NW) {
            const int g = b / NB, blk = b % NB, d = blk * 32 + lane;
            const float v = s_q[g][d / 16][d % 16];
            const float s1 = warp_max(fabsf(v)) * (1.0f / 127.0f);
            const float qh = s1 > 0.f ? rintf(v / s1) : 0.f;
            const float r = v - qh * s1;
            const float s2 = warp_max(fabsf(r)) * (1.0f / 127.0f);
            const float ql = s2 > 0.f ? rintf(r / s2) : 0.f;
            s_qh[g][d] = (int8_t)(int)qh;
            s_ql[g][d] = (int8_t)(int)ql;
            if (lane == 0) { s_qs[g][blk][0] = s1; s_qs[g][blk][1] = s2; }
        }
        __syncthreads();
    }

    float m_run[G], l_run[G], acc[G][4];
#pragma unroll
    for (int g = 0; g < G; g++) {
        m_run[g] = -INFINITY;
        l_run[g] = 0.f;
        acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
    }

    uint32_t phase = 0;
    for (; c < c1; c += NW) {
        mbar_wait(&full_bar[warp], phase);
        phase ^= 1u;
        if (tid == 0 && c == c0) trace_mark(a.trace, 3);
        const uint8_t *tK = tile;
        const uint8_t *tKs = tK + KBYTES;
        const uint8_t *tV = tK + KBYTES + SBYTES;
        const uint8_t *tVs = tV + KBYTES;
        const int pbase = c * ATT_TILE;
        if (a.fuse && c == new_chunk) {
            // the copy raced with (or preceded) the store of the new row: put the row in place here
            const int r = (ctx - 1) - pbase;
            constexpr int ROWB = Q8 ? HD : HD * 2;
            uint8_t *wK = const_cast<uint8_t *>(tK), *wV = const_cast<uint8_t *>(tV);
            for (int i = lane; i < ROWB / 4; i += 32) {
                reinterpret_cast<uint32_t *>(wK + r * ROWB)[i] = reinterpret_cast<const uint32_t *>(s_newrow[0])[i];
                reinterpret_cast<uint32_t *>(wV + r * ROWB)[i] = reinterpret_cast<const uint32_t *>(s_newrow[1])[i];
            }
            if (Q8 && lane < NB) {
                reinterpret_cast<__half *>(const_cast<uint8_t *>(tKs))[r * NB + lane] = s_newsc[0][lane];
                reinterpret_cast<__half *>(const_cast<uint8_t *>(tVs))[r * NB + lane] = s_newsc[1][lane];
            }
            __syncwarp();
        }

        // ---- scores
        if (Q8) {
            // lane = position. A row is HD code bytes; the 16-byte pieces are visited in a per-lane rotated order
            // so that the 8 lanes of an LDS.128 phase hit 8 different bank groups (rows are 128 B or less apart).
            constexpr int C = HD / 16;                 // 16-byte pieces per row
            constexpr int RPL = (C >= 8) ? 1 : 8 / C;  // rows per 128 bytes
            const int rot = lane / RPL;
            float sc[G];
#pragma unroll
            for (int g = 0; g < G; g++) sc[g] = 0.f;
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int ci = (j + rot) & (C - 1);
                const uint4 kw = *reinterpret_cast<const uint4 *>(tK + lane * HD + ci * 16);
                const float ksc = __half2float(reinterpret_cast<const __half *>(tKs)[lane * NB + (ci >> 1)]);
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint4 qh = *reinterpret_cast<const uint4 *>(&s_qh[g][ci * 16]);
                    const uint4 ql = *reinterpret_cast<const uint4 *>(&s_ql[g][ci * 16]);
                    int dh = dp4a_ss(kw.x, qh.x, 0), dl = dp4a_ss(kw.x, ql.x, 0);
                    dh = dp4a_ss(kw.y, qh.y, dh); dl = dp4a_ss(kw.y, ql.y, dl);
                    dh = dp4a_ss(kw.z, qh.z, dh); dl = dp4a_ss(kw.z, ql.z, dl);
                    dh = dp4a_ss(kw.w, qh.w, dh); dl = dp4a_ss(kw.w, ql.w, dl);
                    const float t = fmaf(s_qs[g][ci >> 1][0], (float)dh, s_qs[g][ci >> 1][1] * (float)dl);
                    sc[g] = fmaf(ksc, t, sc[g]);
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) s_p[warp][g][lane] = (pbase + lane < ctx) ? sc[g] : -INFINITY;
        } else {
            // fp16 cache: LPP lanes per position, 16 dims per lane
            const int sub = lane % LPP, pl = lane / LPP;
#pragma unroll 2
            for (int pp = pl; pp < ATT_TILE; pp += PPW) {
                float kf[16];
                const uint4 w0 = *reinterpret_cast<const uint4 *>(tK + (pp * HD + sub * 16) * 2);
                const uint4 w1 = *reinterpret_cast<const uint4 *>(tK + (pp * HD + sub * 16) * 2 + 16);
                const __half2 *h0 = reinterpret_cast<const __half2 *>(&w0);
                const __half2 *h1 = reinterpret_cast<const __half2 *>(&w1);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
                    kf[2 * j] = f0.x; kf[2 * j + 1] = f0.y;
                    kf[8 + 2 * j] = f1.x; kf[8 + 2 * j + 1] = f1.y;
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const float4 *qv = reinterpret_cast<const float4 *>(&s_q[g][sub][0]);
                    float s0 = 0.f, s1 = 0.f;   // two chains: halve the dependent-FMA latency
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const float4 qa = qv[j], qb = qv[j + 1];
                        s0 = fmaf(kf[4 * j], qa.x, s0);
                        s1 = fmaf(kf[4 * j + 4], qb.x, s1);
                        s0 = fmaf(kf[4 * j + 1], qa.y, s0);
                        s1 = fmaf(kf[4 * j + 5], qb.y, s1);
                        s0 = fmaf(kf[4 * j + 2], qa.z, s0);
                        s1 = fmaf(kf[4 * j + 6], qb.z, s1);
                        s0 = fmaf(kf[4 * j + 3], qa.w, s0);
                        s1 = fmaf(kf[4 * j + 7], qb.w, s1);
                    }
                    float sc = s0 + s1;
#pragma unroll
                    for (int o = LPP / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
                    if (sub == 0) s_p[warp][g][pp] = (pbase + pp < ctx) ? sc : -INFINITY;
                }
            }
        }
        __syncwarp();
        if (tid == 0 && c == c0) trace_mark(a.trace, 5);

        // ---- online softmax for the G heads (whole warp per head)
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float v0 = s_p[warp][g][lane];
            const float cm = warp_max(v0);
            const float m_new = fmaxf(m_run[g], cm);
            const float e0 = (v0 == -INFINITY) ? 0.f : __expf(v0 - m_new);
            const float cs = warp_sum(e0);
            s_p[warp][g][lane] = e0;
            const float corr = (m_run[g] == -INFINITY) ? 0.f : __expf(m_run[g] - m_new);
            l_run[g] = l_run[g] * corr + cs;
            m_run[g] = m_new;
            acc[g][0] *= corr; acc[g][1] *= corr; acc[g][2] *= corr; acc[g][3] *= corr;
        }
        __syncwarp();
        if (tid == 0 && c == c0) trace_mark(a.trace, 6);

        // ---- P.V: lane owns 4 dims (one code word), loop over the tile's positions
        if (lane < HD / 4) {
            const int npos = min(ATT_TILE, ctx - pbase);
#pragma unroll 4
            for (int pp = 0; pp < npos; pp++) {
                float vf[4], vsc = 1.f;
                if (Q8) {
                    const uint32_t w = s8bias(reinterpret_cast<const uint32_t *>(tV + pp * HD)[lane]);
                    vsc = __half2float(reinterpret_cast<const __half *>(tVs)[pp * NB + lane / 8]);
#pragma unroll
                    for (int k = 0; k < 4; k++) vf[k] = s8f(w, k);
                } else {
                    const uint2 w = reinterpret_cast<const uint2 *>(tV + pp * HD * 2)[lane];
                    float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&w.x));
                    float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
                    vf[0] = f0.x; vf[1] = f0.y; vf[2] = f1.x; vf[3] = f1.y;
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const float p = s_p[warp][g][pp] * vsc;   // the block scale rides on the probability
                    acc[g][0] = fmaf(p, vf[0], acc[g][0]);
                    acc[g][1] = fmaf(p, vf[1], acc[g][1]);
                    acc[g][2] = fmaf(p, vf[2], acc[g][2]);
                    acc[g][3] = fmaf(p, vf[3], acc[g][3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && c + NW < c1) issue(c + NW);   // the buffer is free again
    }
    if (tid == 0) trace_mark(a.trace, 4);

    // ---- merge the 4 warps (log-sum-exp) through shared memory, write the CTA's partial
    __syncthreads();                                   // every tile buffer is idle now: reuse as scratch
    float *red = reinterpret_cast<float *>(smem);      // [NW][G][HD]
    if (lane < HD / 4) {
#pragma unroll
        for (int g = 0; g < G; g++)
            *reinterpret_cast<float4 *>(&red[((warp * G) + g) * HD + lane * 4]) =
                make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
    }
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < G; g++) { s_ml[warp][g][0] = m_run[g]; s_ml[warp][g][1] = l_run[g]; }
    }
    __syncthreads();
    for (int i = tid; i < G * HD; i += ATT_THREADS) {
        const int g = i / HD, d = i % HD;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; w++) M = fmaxf(M, s_ml[w][g][0]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const float mw = s_ml[w][g][0];
            const float wt = (mw == -INFINITY) ? 0.f : __expf(mw - M);
            num = fmaf(wt, red[((w * G) + g) * HD + d], num);
            den = fmaf(wt, s_ml[w][g][1], den);
        }
        float *o;
        if (a.cluster) {
            // straight into the shared memory of the cluster's first CTA (distributed shared memory)
            float *lead = cg::this_cluster().map_shared_rank(reinterpret_cast<float *>(smem + a.part_off), 0);
            o = lead + ((int64_t)g * a.nsplit + split) * (HD + 2);
        } else {
            o = a.ws + ((int64_t)(kvh * G + g) * a.nsplit + split) * (HD + 2);
        }
        o[2 + d] = num;
        if (d == 0) { o[0] = M; o[1] = den; }
    }
    if (a.cluster) {
        // Every CTA of the head has delivered once the cluster barrier completes (release/acquire at cluster scope):
        // no fence, no global round trip, no atomic. The first CTA finishes the log-sum-exp merge from its own
        // shared memory.
        cg::this_cluster().sync();
        if (split != 0) return;
        const float *pbase = reinterpret_cast<const float *>(smem + a.part_off);
        float *wts = reinterpret_cast<float *>(smem);      // [G][nsplit] normalised weights (tile buffers are idle)
        for (int g = warp; g < G; g += NW) {
            const float *base = pbase + (int64_t)g * a.nsplit * (HD + 2);
            float M = -INFINITY;
            for (int sp = lane; sp < a.nsplit; sp += 32) M = fmaxf(M, base[sp * (HD + 2)]);
            M = warp_max(M);
            float den = 0.f;
            for (int sp = lane; sp < a.nsplit; sp += 32) {
                const float ms = base[sp * (HD + 2)];
                const float wt = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                wts[g * a.nsplit + sp] = wt;
                den = fmaf(wt, base[sp * (HD + 2) + 1], den);
            }
            den = warp_sum(den);
            const float inv = 1.0f / den;
            for (int sp = lane; sp < a.nsplit; sp += 32) wts[g * a.nsplit + sp] *= inv;
        }
        __syncthreads();
        for (int i = tid; i < G * HD; i += ATT_THREADS) {
            const int g = i / HD, d = i % HD;
            const float *base = pbase + (int64_t)g * a.nsplit * (HD + 2) + 2 + d;
            float num = 0.f;
#pragma unroll 4
            for (int sp = 0; sp < a.nsplit; sp++) num = fmaf(wts[g * a.nsplit + sp], base[(int64_t)sp * (HD + 2)], num);
            a.out[(int64_t)(kvh * G + g) * HD + d] = __float2half_rn(num);
        }
        return;
    }
    if (a.nsplit == 1) {
        // single CTA per head: normalise right here
        __syncthreads();
        for (int i = tid; i < G * HD; i += ATT_THREADS) {
            const int g = i / HD, d = i % HD;
            const float *o = a.ws + ((int64_t)(kvh * G + g) * a.nsplit) * (HD + 2);
            a.out[(int64_t)(kvh * G + g) * HD + d] = __float2half_rn(o[2 + d] / o[1]);
        }
        return;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) trace_mark(a.trace, 7);
    if (tid == 0) {
        unsigned int prev = atomicAdd(&a.counters[kvh], 1u);
        s_last = (prev == (unsigned)a.nsplit - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u) return;
    __threadfence();
    // ---- the last CTA of this kv head merges the CTAs' partials
    float *wts = reinterpret_cast<float *>(smem);      // [G][nsplit] normalised weights
    for (int g = warp; g < G; g += NW) {
        const float *base = a.ws + (int64_t)(kvh * G + g) * a.nsplit * (HD + 2);
        float M = -INFINITY;
        for (int sp = lane; sp < a.nsplit; sp += 32) M = fmaxf(M, base[sp * (HD + 2)]);
        M = warp_max(M);
        float den = 0.f;
        for (int sp = lane; sp < a.nsplit; sp += 32) {
            const float ms = base[sp * (HD + 2)];
            const float wt = (ms == -INFINITY) ? 0.f : __expf(ms - M);
            wts[g * a.nsplit + sp] = wt;
            den = fmaf(wt, base[sp * (HD + 2) + 1], den);
        }
        den = warp_sum(den);
        const float inv = 1.0f / den;
        for (int sp = lane; sp < a.nsplit; sp += 32) wts[g * a.nsplit + sp] *= inv;
    }
    __syncthreads();
    for (int i = tid; i < G * HD; i += ATT_THREADS) {
        const int g = i / HD, d = i % HD;
        const float *base = a.ws + (int64_t)(kvh * G + g) * a.nsplit * (HD + 2) + 2 + d;
        float num = 0.f;
#pragma unroll 4
        for (int sp = 0; sp < a.nsplit; sp++) num = fmaf(wts[g * a.nsplit + sp], base[(int64_t)sp * (HD + 2)], num);
        a.out[(int64_t)(kvh * G + g) * HD + d] = __float2half_rn(num);
    }
    if (tid == 0) a.counters[kvh] = 0u;  // ready for the next launch
}

template <int HD, int G>
static int attn_launch_t(const AttnArgs &a_in, int kv_dtype, bool pdl, cudaStream_t st) {
    AttnArgs a = a_in;
    const bool q8 = kv_dtype == DT_Q8_B32T2;
    const int nb = HD / 32;
    const size_t tile = q8 ? 2 * (ATT_TILE * HD + ATT_TILE * nb * 2) : 2 * (ATT_TILE * HD * 2);
    size_t smem = (ATT_THREADS / 32) * tile;
    const size_t red = (size_t)(ATT_THREADS / 32) * G * HD * sizeof(float);
    const size_t wts = (size_t)G * a.nsplit * sizeof(float);
    if (smem < red) smem = red;
    if (smem < wts) smem = wts;
    a.part_off = 0;
    if (a.cluster) {   // the leader's partials live behind the tile buffers: peers write them while it still scans
        smem = (smem + 127) / 128 * 128;
        a.part_off = (int)smem;
        smem += (size_t)a.nsplit * G * (HD + 2) * sizeof(float);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.nsplit, a.kv_heads);
    cfg.blockDim = dim3(ATT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (a.cluster) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = a.nsplit; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        na++;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    static size_t attr_smem[64][2] = {};   // per <HD, G> instantiation and device: the opt-in is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (q8) {
        auto kern = attn_decode_kernel<HD, G, true>;
        if (attr_smem[dev][1] < smem) { IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem[dev][1] = smem; }
        IFB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        auto kern = attn_decode_kernel<HD, G, false>;
        if (attr_smem[dev][0] < smem) { IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem[dev][0] = smem; }
        IFB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    }
    return IFB_OK;
}

template <int HD>
static int attn_launch_hd(const AttnArgs &a, int group, int kv_dtype, bool pdl, cudaStream_t st) {
    switch (group) {
        case 1: return attn_launch_t<HD, 1>(a, kv_dtype, pdl, st);
        case 2: return attn_launch_t<HD, 2>(a, kv_dtype, pdl, st);
        case 4: return attn_launch_t<HD, 4>(a, kv_dtype, pdl, st);
        case 8: return attn_launch_t<HD, 8>(a, kv_dtype, pdl, st);
        default:
            set_last_error("attention: heads/kv_heads = %d not in {1,2,4,8}", group);
            return IFB_ERR_ARG;
    }
}

int attn_decode(const AttnDecode &p, cudaStream_t st) {
    IFB_CHECK_ARG(p.kv_dtype == DT_F16 || p.kv_dtype == DT_Q8_B32T2, "kv cache type %d", p.kv_dtype);
    IFB_CHECK_ARG(p.kv_heads > 0 && p.heads % p.kv_heads == 0, "attention: heads %d / kv_heads %d", p.heads, p.kv_heads);
    IFB_CHECK_ARG(p.pos_dev != nullptr || (p.ctx_len >= 1 && p.ctx_len <= p.max_ctx), "attention: ctx_len %d", p.ctx_len);
    AttnArgs a;
    a.heads = p.heads; a.kv_heads = p.kv_heads;
    a.max_ctx_pad = round_up(p.max_ctx, ATT_CHUNK);
    a.ctx_len = p.ctx_len;
    a.softmax_scale = 1.0f / sqrtf((float)p.head_dim);   // alpha*kq_scale (inference_worker.cc:1666, softmax s_scale)
    a.out = (__half *)p.out; a.q = (const __half *)p.q;
    a.kcache = (const uint8_t *)p.kcache; a.vcache = (const uint8_t *)p.vcache;
    a.counters = reinterpret_cast<unsigned int *>(p.workspace);
    a.ws = reinterpret_cast<float *>((uint8_t *)p.workspace + 4096);
    a.pos_dev = p.pos_dev;
    a.trace = (unsigned long long *)p.trace;
    a.k_new = (const __half *)p.k_new; a.v_new = (const __half *)p.v_new;
    a.fuse = p.k_new != nullptr ? 1 : 0;
    IFB_CHECK_ARG((p.k_new == nullptr) == (p.v_new == nullptr), "attention: fused append needs both the new K and V rows");
    a.rope = p.rope; a.rope_order = p.rope_order; a.rope_dims = p.rope_dims; a.rope_theta = p.rope_theta;
    IFB_CHECK_ARG(!a.fuse || !a.rope || ((p.rope_order == 0 || p.rope_order == 2) && p.rope_dims % 2 == 0 && p.rope_dims <= p.head_dim),
                  "attention: rope order %d / dims %d", p.rope_order, p.rope_dims);
    // split count: fill ~2 CTAs per SM; with a device-side position (graph replay) size for max_ctx
    const int ctx_for_split = p.pos_dev != nullptr ? p.max_ctx : p.ctx_len;
    const int nchunks = (ctx_for_split + ATT_TILE - 1) / ATT_TILE;
    // one tile per warp while that still fills the machine: CTAs per head = ceil(tiles / 4),
    // capped at ~4 CTAs per SM overall (longer contexts: several tiles per warp)
    int nsplit = (nchunks + (ATT_THREADS / 32) - 1) / (ATT_THREADS / 32);
    const int cap = (4 * sm_count() + p.kv_heads - 1) / p.kv_heads;
    if (nsplit > cap) nsplit = cap;
    if (nsplit > 4 * 160) nsplit = 4 * 160;
    if (nsplit < 1) nsplit = 1;
    // the splits of a head merge through distributed shared memory when they fit one portable cluster (<= 8 CTAs);
    // longer contexts get more tiles per warp instead of more CTAs
    static int no_cluster = -1;
    if (no_cluster < 0) { const char *e = getenv("IFB_ATTN_NO_CLUSTER"); no_cluster = (e && atoi(e)) ? 1 : 0; }
    a.cluster = 0;
    if (!no_cluster && nsplit > 1) {
        if (nsplit > 8) nsplit = 8;
        a.cluster = 1;
    }
    a.nsplit = nsplit;
    const int group = p.heads / p.kv_heads;
    switch (p.head_dim) {
        case 32: return attn_launch_hd<32>(a, group, p.kv_dtype, p.pdl, st);
        case 64: return attn_launch_hd<64>(a, group, p.kv_dtype, p.pdl, st);
        case 128: return attn_launch_hd<128>(a, group, p.kv_dtype, p.pdl, st);
        default:
            set_last_error("attention: head_dim %d not supported (32/64/128)", p.head_dim);
            return IFB_ERR_ARG;
    }
}

}  // namespace ifb
