// Prompt processing (rows > 1): causal attention of T query tokens over the KV cache, and the helpers of the
// multi-token layer pass of ifb_model_prefill.
//
// Replaces ProcessGpuLayer_Attention for q_tokens > 1 (reference src/transformer/inference_worker.cc:1130-1320,
// :1639-1724; tensor_mul.cu:232-337 Gemm_Alg2; unary_tensor_opr.h:481 SoftMax): the reference dequantizes the whole K
// and V prefix, repeats the kv heads, runs two batched GEMMs through a [heads][T][ctx] score tensor and a causal softmax
// over it -- and refuses prompts beyond its scratch size (:2612-2618). Here the score tensor never exists: one CTA owns
// 64 queries of one head, walks the cache 64 positions at a time (tiles dequantized into shared memory from the cache as
// stored), S = Q.K^T and O += P.V on the tensor cores (mma.sync tiles through the WMMA API, fp16 in / fp32 out), online
// softmax in fp32 between them. Any prompt length the cache holds works.
#include <mma.h>

#include "ifb_formats.cuh"
#include "ifb_kernels.h"

namespace ifb {

using namespace nvcuda;

constexpr int PF_BQ = 64;        // queries per CTA (4 warps x 16)
constexpr int PF_BK = 64;        // cache positions per tile
constexpr int PF_THREADS = 128;

struct PrefillArgs {
    const __half *q;             // [T][heads*HD], RoPE applied
    __half *out;                 // [T][heads*HD]
    const uint8_t *kcache, *vcache;
    int T, pos0, heads, kv_heads, max_ctx_pad;
    float scale;                 // 1/sqrt(head_dim)
};

template <int HD, bool Q8>
__global__ void __launch_bounds__(PF_THREADS) attn_prefill_kernel(const PrefillArgs a) {
    constexpr int NB = HD / 32;
    constexpr int LDH = HD + 8;          // half tiles: row stride (multiple of 8 halves, off the 128-byte bank period)
    constexpr int LDS_ = PF_BK + 4;      // score rows (floats)
    constexpr int LDP = PF_BK + 8;       // probability rows (halves)
    constexpr int LDO = HD + 4;          // output accumulator rows (floats)
    extern __shared__ __align__(128) uint8_t smem[];
    __half *Qs = reinterpret_cast<__half *>(smem);                       // [64][LDH]
    __half *Ks = Qs + PF_BQ * LDH;                                       // [64][LDH]
    __half *Vs = Ks + PF_BK * LDH;                                       // [64][LDH]
    float *Ss = reinterpret_cast<float *>(Vs + PF_BK * LDH);             // [4][16][LDS_]
    float *Os = Ss + 4 * 16 * LDS_;                                      // [4][16][LDO]
    __half *Ps = reinterpret_cast<__half *>(Os + 4 * 16 * LDO);          // [4][16][LDP]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y, kvh = h / (a.heads / a.kv_heads);
    const int q0 = blockIdx.x * PF_BQ;
    const int qdim = a.heads * HD;

    // ---- the CTA's 64 queries of head h (zeros beyond T)
    for (int i = tid; i < PF_BQ * (HD / 8); i += PF_THREADS) {
        const int r = i / (HD / 8), c = i % (HD / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + r < a.T) v = *reinterpret_cast<const uint4 *>(a.q + (size_t)(q0 + r) * qdim + h * HD + c * 8);
        *reinterpret_cast<uint4 *>(Qs + r * LDH + c * 8) = v;
    }
    float *Ow = Os + warp * 16 * LDO;
    for (int i = lane; i < 16 * LDO; i += 32) Ow[i] = 0.f;
    // softmax state of row (lane >> 1) of this warp; the two lanes of a row keep identical copies
    float m_run = -INFINITY, l_run = 0.f;
    const int row = lane >> 1, half_ = lane & 1;
    const int qpos = a.pos0 + q0 + warp * 16 + row;            // absolute position of this lane's query
    const bool q_ok = q0 + warp * 16 + row < a.T;

    const int64_t head_rows = (int64_t)kvh * a.max_ctx_pad;
    const uint8_t *kc = a.kcache + (Q8 ? head_rows * HD : head_rows * HD * 2);
    const uint8_t *vc = a.vcache + (Q8 ? head_rows * HD : head_rows * HD * 2);
    const __half *ksc = reinterpret_cast<const __half *>(a.kcache + (int64_t)a.kv_heads * a.max_ctx_pad * HD) + head_rows * NB;
    const __half *vsc = reinterpret_cast<const __half *>(a.vcache + (int64_t)a.kv_heads * a.max_ctx_pad * HD) + head_rows * NB;

    const int last_q = min(q0 + PF_BQ, a.T) - 1;
    const int kv_end = a.pos0 + last_q + 1;                    // positions [0, kv_end) can be attended by this CTA
    for (int p0 = 0; p0 < kv_end; p0 += PF_BK) {
        __syncthreads();                                       // the previous tile's K/V are no longer read
        // ---- stage the K and V tile as fp16 (Q8_B32T2: code * scale; the cache is padded to 64 positions)
        for (int i = tid; i < PF_BK * (HD / 16); i += PF_THREADS) {
            const int r = i / (HD / 16), c = i % (HD / 16);    // position in the tile, 16-dim chunk
            const int64_t prow = p0 + r;
            if (Q8) {
                const uint4 kw = *reinterpret_cast<const uint4 *>(kc + prow * HD + c * 16);
                const uint4 vw = *reinterpret_cast<const uint4 *>(vc + prow * HD + c * 16);
                const float ks = __half2float(ksc[prow * NB + (c >> 1)]), vs = __half2float(vsc[prow * NB + (c >> 1)]);
                const uint32_t kk[4] = {kw.x, kw.y, kw.z, kw.w}, vv[4] = {vw.x, vw.y, vw.z, vw.w};
                __half2 ko[8], vo[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int k0 = (int)(int8_t)(kk[j] & 0xff), k1 = (int)(int8_t)((kk[j] >> 8) & 0xff);
                    const int k2 = (int)(int8_t)((kk[j] >> 16) & 0xff), k3 = (int)(int8_t)(kk[j] >> 24);
                    const int v0 = (int)(int8_t)(vv[j] & 0xff), v1 = (int)(int8_t)((vv[j] >> 8) & 0xff);
                    const int v2 = (int)(int8_t)((vv[j] >> 16) & 0xff), v3 = (int)(int8_t)(vv[j] >> 24);
                    // GetKRows / GetVRows dequantize to half: value = code * scale (kv_cache.cc:185-249)
                    ko[2 * j] = __floats2half2_rn((float)k0 * ks, (float)k1 * ks);
                    ko[2 * j + 1] = __floats2half2_rn((float)k2 * ks, (float)k3 * ks);
                    vo[2 * j] = __floats2half2_rn((float)v0 * vs, (float)v1 * vs);
                    vo[2 * j + 1] = __floats2half2_rn((float)v2 * vs, (float)v3 * vs);
                }
                *reinterpret_cast<uint4 *>(Ks + r * LDH + c * 16) = *reinterpret_cast<const uint4 *>(&ko[0]);
                *reinterpret_cast<uint4 *>(Ks + r * LDH + c * 16 + 8) = *reinterpret_cast<const uint4 *>(&ko[4]);
                *reinterpret_cast<uint4 *>(Vs + r * LDH + c * 16) = *reinterpret_cast<const uint4 *>(&vo[0]);
                *reinterpret_cast<uint4 *>(Vs + r * LDH + c * 16 + 8) = *reinterpret_cast<const uint4 *>(&vo[4]);
            } else {
                const uint4 *kp = reinterpret_cast<const uint4 *>(kc + (prow * HD + c * 16) * 2);
                const uint4 *vp = reinterpret_cast<const uint4 *>(vc + (prow * HD + c * 16) * 2);
                *reinterpret_cast<uint4 *>(Ks + r * LDH + c * 16) = kp[0];
                *reinterpret_cast<uint4 *>(Ks + r * LDH + c * 16 + 8) = kp[1];
                *reinterpret_cast<uint4 *>(Vs + r * LDH + c * 16) = vp[0];
                *reinterpret_cast<uint4 *>(Vs + r * LDH + c * 16 + 8) = vp[1];
            }
        }
        __syncthreads();
        if (p0 > a.pos0 + q0 + warp * 16 + 15) continue;       // every position of the tile is in this warp's future

        // ---- S = Q_w . K^T : 16 x 64, fp32
        float *Sw = Ss + warp * 16 * LDS_;
        {
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[PF_BK / 16];
#pragma unroll
            for (int n = 0; n < PF_BK / 16; n++) wmma::fill_fragment(acc[n], 0.f);
#pragma unroll
            for (int k = 0; k < HD / 16; k++) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> fa;
                wmma::load_matrix_sync(fa, Qs + (warp * 16) * LDH + k * 16, LDH);
#pragma unroll
                for (int n = 0; n < PF_BK / 16; n++) {
                    wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> fb;     // K^T: (k, n) at Ks[n][k]
                    wmma::load_matrix_sync(fb, Ks + (n * 16) * LDH + k * 16, LDH);
                    wmma::mma_sync(acc[n], fa, fb, acc[n]);
                }
            }
#pragma unroll
            for (int n = 0; n < PF_BK / 16; n++) wmma::store_matrix_sync(Sw + n * 16, acc[n], LDS_, wmma::mem_row_major);
        }
        __syncwarp();
        // ---- online softmax of the warp's 16 rows (lane = row / half of the 64 columns), fp32
        {
            float sv[32];
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int col = half_ * 32 + c;
                float s = Sw[row * LDS_ + col] * a.scale;
                if (!q_ok || p0 + col > qpos) s = -INFINITY;           // causal mask (and rows beyond the prompt)
                sv[c] = s;
                tmax = fmaxf(tmax, s);
            }
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
            const float m_new = fmaxf(m_run, tmax);
            const float alpha = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
            float rs = 0.f;
            __half *Pw = Ps + warp * 16 * LDP + row * LDP + half_ * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float e0 = (sv[c] == -INFINITY) ? 0.f : __expf(sv[c] - m_new);
                const float e1 = (sv[c + 1] == -INFINITY) ? 0.f : __expf(sv[c + 1] - m_new);
                const __half2 hp = __floats2half2_rn(e0, e1);
                *reinterpret_cast<__half2 *>(Pw + c) = hp;
                const float2 back = __half22float2(hp);                // the sum of what P.V will actually use
                rs += back.x + back.y;
            }
            rs += __shfl_xor_sync(0xffffffffu, rs, 1);
            l_run = l_run * alpha + rs;
            m_run = m_new;
            // rescale this row's accumulator (each of the two lanes takes half of the row)
            float *orow = Ow + row * LDO + half_ * (HD / 2);
#pragma unroll 8
            for (int d = 0; d < HD / 2; d++) orow[d] *= alpha;
        }
        __syncwarp();
        // ---- O_w += P . V : 16 x HD
#pragma unroll
        for (int n = 0; n < HD / 16; n++) {
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
            wmma::load_matrix_sync(acc, Ow + n * 16, LDO, wmma::mem_row_major);
#pragma unroll
            for (int k = 0; k < PF_BK / 16; k++) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> fp;
                wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> fv;
                wmma::load_matrix_sync(fp, Ps + warp * 16 * LDP + k * 16, LDP);
                wmma::load_matrix_sync(fv, Vs + (k * 16) * LDH + n * 16, LDH);
                wmma::mma_sync(acc, fp, fv, acc);
            }
            wmma::store_matrix_sync(Ow + n * 16, acc, LDO, wmma::mem_row_major);
        }
        __syncwarp();
    }
    // ---- out = O / l, as half (the attention output tensor is half)
    if (q_ok) {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        const float *orow = Ow + row * LDO + half_ * (HD / 2);
        __half *dst = a.out + (size_t)(q0 + warp * 16 + row) * qdim + h * HD + half_ * (HD / 2);
#pragma unroll 8
        for (int d = 0; d < HD / 2; d += 2) *reinterpret_cast<__half2 *>(dst + d) = __floats2half2_rn(orow[d] * inv, orow[d + 1] * inv);
    }
}

template <int HD, bool Q8>
static int launch_prefill(const PrefillArgs &a, cudaStream_t st) {
    auto kern = attn_prefill_kernel<HD, Q8>;
    const size_t smem = (size_t)(PF_BQ + 2 * PF_BK) * (HD + 8) * 2 + (size_t)4 * 16 * (PF_BK + 4) * 4 + (size_t)4 * 16 * (HD + 4) * 4 +
                        (size_t)4 * 16 * (PF_BK + 8) * 2;
    static int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev = dev;
    }
    dim3 grid((a.T + PF_BQ - 1) / PF_BQ, a.heads);
    kern<<<grid, PF_THREADS, smem, st>>>(a);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

// causal attention of T query tokens at positions pos0..pos0+T-1 (their K/V rows already appended) over cache positions
// 0..pos0+T-1. q / out: fp16 [T][heads*head_dim], q with RoPE applied.
int attn_prefill(int kv_dtype, void *out_f16, const void *q_f16, const void *kcache, const void *vcache, int tokens, int pos0,
                 int heads, int kv_heads, int head_dim, int max_ctx, cudaStream_t st) {
    IFB_CHECK_ARG(kv_dtype == DT_F16 || kv_dtype == DT_Q8_B32T2, "prefill attention: kv cache type %d", kv_dtype);
    IFB_CHECK_ARG(tokens >= 1 && pos0 >= 0 && pos0 + tokens <= max_ctx, "prefill attention: positions %d..%d outside max_context_len %d", pos0, pos0 + tokens, max_ctx);
    IFB_CHECK_ARG(kv_heads > 0 && heads % kv_heads == 0, "prefill attention: heads %d / kv_heads %d", heads, kv_heads);
    IFB_CHECK_ARG(head_dim == 64 || head_dim == 128, "prefill attention: head_dim %d (64 / 128)", head_dim);
    PrefillArgs a;
    a.q = (const __half *)q_f16; a.out = (__half *)out_f16; a.kcache = (const uint8_t *)kcache; a.vcache = (const uint8_t *)vcache;
    a.T = tokens; a.pos0 = pos0; a.heads = heads; a.kv_heads = kv_heads;
    a.max_ctx_pad = (max_ctx + 63) / 64 * 64;
    a.scale = 1.0f / sqrtf((float)head_dim);
    const bool q8 = kv_dtype == DT_Q8_B32T2;
    if (head_dim == 128) return q8 ? launch_prefill<128, true>(a, st) : launch_prefill<128, false>(a, st);
    return q8 ? launch_prefill<64, true>(a, st) : launch_prefill<64, false>(a, st);
}

// ---------------------------------------------------------------- helpers of the multi-token layer pass
// embedding rows of T tokens (GetEmbdTensor, inference_engine.cc:1324-1350)
__global__ void embed_rows_kernel(__half *dst, const __half *table, const int32_t *tokens, int cols, int vocab) {
    int tok = tokens[blockIdx.x];
    tok = max(0, min(tok, vocab - 1));
    const uint4 *src = reinterpret_cast<const uint4 *>(table + (size_t)tok * cols);
    uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)blockIdx.x * cols);
    for (int i = threadIdx.x; i < cols / 8; i += blockDim.x) d[i] = src[i];
}
int embed_rows(void *dst_f16, const void *table_f16, const int32_t *tokens_dev, int n, int64_t cols, int vocab, cudaStream_t st) {
    IFB_CHECK_ARG(n >= 1 && cols % 8 == 0, "embed_rows: bad shape");
    embed_rows_kernel<<<n, 128, 0, st>>>((__half *)dst_f16, (const __half *)table_f16, tokens_dev, (int)cols, vocab);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

}  // namespace ifb
