// Decode GEMV for Inferflow block-quantized weights, sm_100a.
//
// Replaces TensorMul::Gemv_AX -> Gemv_AX_Alg3 -> Gemv_AX_QuantQ8 / Gemv_AX_QuantHalf / fp16
// (reference src/tensor/tensor_mul.cu:97,766,714,658; kernels src/kernels/gemv.h:470-1709)
// and, in its fused form, the RmsNorm -> Quantize(Q8_B32T2) -> GEMV(s) -> SiLU*Mul / +residual
// chains of GpuInferenceWorker (src/transformer/inference_worker.cc:1038,1442-1444,
// 1555-1571,1726-1900).
//
// Design (HBM-bound streaming kernel; DESIGN.md "gemv"):
//  * persistent CTAs, 1 producer warp + 8 consumer warps;
//  * the weight matrix is streamed in the reference's own AoS block layout: a "unit" is R
//    whole rows = one contiguous byte range, moved global->shared by ONE 1-D TMA bulk copy
//    (cp.async.bulk, SASS UBLKCP) into a ring of shared-memory slots guarded by full/empty
//    mbarriers -- the producer runs ahead of the consumers by the ring depth and, under PDL,
//    starts streaming BEFORE the previous kernel has finished (weights do not depend on it);
//  * consumers: one warp per row, lanes stride over the row's blocks (20/36/52-byte block
//    strides are odd multiples of 4 bytes -> conflict-free LDS.32), integer dot by dp4a
//    against the int8 activation staged once per CTA, fp32 combine, warp-shuffle reduce;
//  * the activation prologue (optional RMSNorm, Q8_B32T2 quantization identical to
//    Tensor_QuantizeQ8_B32T2_Alg2_Kernel, kernels/tensor_quant.h:45-82) runs in every CTA
//    from the fp16 vector, overlapped with the first weight copies.
#include "ifb_formats.cuh"
#include "ifb_kernels.h"

namespace ifb {

constexpr int GEMV_CWARPS = 8;                       // consumer warps
constexpr int GEMV_THREADS = (GEMV_CWARPS + 1) * 32; // + producer warp
constexpr int GEMV_MAX_SLOTS = 16;
constexpr int BAR_CONSUMERS = 1;                     // named barrier id for the 8 consumer warps

struct GemvArgs {
    const uint8_t *w[3];
    int rows[3];
    int unit0[3];  // first unit index of each segment
    int nseg;
    const uint8_t *w_gate;
    int cx, nblk, row_bytes;
    int R;           // rows per unit/slot
    int wpg;         // consumer warps per group (= min(R, 8))
    int nslots, slot_stride;
    int units;
    const __half *x;
    const uint8_t *xq;  // pre-quantized activation (Q8_B32T2 blocks) or null
    const __half *norm_w;
    float eps, norm_base;
    const __half *residual, *bias;
    __half *y;
    int w_dtype;
};

// ---------------------------------------------------------------------------------------
// activation staging
// ---------------------------------------------------------------------------------------
// byte position of element e (0..31) of an activation block inside its 32 staged bytes
template <bool XSPLIT>
__device__ __forceinline__ int x_pos(int e) {
    if (!XSPLIT) return e;
    int g = e >> 3, k = e & 7;
    return g * 8 + (k & 1) * 4 + (k >> 1);
}

// RMS scale of the whole vector, computed by the 256 consumer threads (128-bit loads).
// Tensor_RmsNorm_Kernel (unary_tensor_opr.h:217): scale = 1/sqrt(mean(x^2) + eps)
__device__ __forceinline__ float rms_scale_256(const __half *x, int cx, float eps, float *red,
                                               int tid) {
    const uint4 *x4 = reinterpret_cast<const uint4 *>(x);
    const int ng = cx >> 3;
    float s = 0.f;
#pragma unroll 4
    for (int g = tid; g < ng; g += GEMV_CWARPS * 32) {
        const uint4 v = x4[g];
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 f = __half22float2(h[k]);
            s = fmaf(f.x, f.x, s);
            s = fmaf(f.y, f.y, s);
        }
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    named_bar_sync(BAR_CONSUMERS, GEMV_CWARPS * 32);
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < GEMV_CWARPS; i++) t += red[i];
    const float mean = t / (float)cx;
    return 1.0f / sqrtf(mean + eps);
}

// Stage x as Q8_B32T2 (codes + fp16-rounded scale + scale*sum(q)) into shared memory.
// The arithmetic per block of 32 is the reference's alg 2 (tensor_quant.h:45-82): scale =
// absmax/127, q = roundf(v/scale), scale <= 1e-6 -> 0. Mapping: one thread = 8 consecutive
// elements (ONE 128-bit load), four neighbouring lanes = one block; all of a thread's loads are
// issued before any is used, so the prologue costs one memory round trip, not one per block.
template <bool XSPLIT>
__device__ __forceinline__ void stage_x_q8(const GemvArgs &a, uint8_t *xlo, uint8_t *xhi,
                                           float2 *xs, float *red, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int nxb = a.cx >> 5;
    if (a.xq != nullptr) {  // already quantized by the caller (TensorMul::Gemv_AX with a Q8 X)
        for (int b = warp; b < nxb; b += GEMV_CWARPS) {
            const uint8_t *blk = a.xq + (size_t)b * 34;
            int q = (int)(int8_t)blk[2 + lane];
            int p = x_pos<XSPLIT>(lane);
            (p < 16 ? xlo + b * 16 + p : xhi + b * 16 + (p - 16))[0] = (uint8_t)q;
            int qs = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) {
                float s = h2f(*(const uint16_t *)blk);
                xs[b] = make_float2(s, s * (float)qs);
            }
        }
        return;
    }
    const bool norm = a.norm_w != nullptr;
    float rs = 1.f;
    if (norm) rs = rms_scale_256(a.x, a.cx, a.eps, red, tid);
    const uint4 *x4 = reinterpret_cast<const uint4 *>(a.x);
    const uint4 *w4 = reinterpret_cast<const uint4 *>(a.norm_w);
    const int ng = a.cx >> 3;                 // groups of 8 elements; ng % 4 == 0
    constexpr int KPT = 4;                    // groups per thread per round
    for (int g0 = 0; g0 < ng; g0 += GEMV_CWARPS * 32 * KPT) {
        uint4 xv[KPT], wv[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * GEMV_CWARPS * 32 + tid;
            xv[k] = make_uint4(0, 0, 0, 0);
            wv[k] = make_uint4(0, 0, 0, 0);
            if (g < ng) {
                xv[k] = x4[g];
                if (norm) wv[k] = w4[g];
            }
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * GEMV_CWARPS * 32 + tid;   // a quad of lanes shares one block
            float v[8];
            {
                const __half2 *h = reinterpret_cast<const __half2 *>(&xv[k]);
                const __half2 *hw = reinterpret_cast<const __half2 *>(&wv[k]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 f = __half22float2(h[i]);
                    if (norm) {  // v = x*scale*(multi_base + w), stored as half by the reference's RmsNorm
                        const float2 fw = __half22float2(hw[i]);
                        f.x = round_h(f.x * rs * (a.norm_base + fw.x));
                        f.y = round_h(f.y * rs * (a.norm_base + fw.y));
                    }
                    v[2 * i] = f.x;
                    v[2 * i + 1] = f.y;
                }
            }
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) m = fmaxf(m, fabsf(v[i]));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            const float scale = m / 127.0f;
            int q[8], qs = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int t = scale <= 0.000001f ? 0 : (int)roundf(v[i] / scale);
                q[i] = min(max(t, -128), 127);
                qs += q[i];
            }
            qs += __shfl_xor_sync(0xffffffffu, qs, 1);
            qs += __shfl_xor_sync(0xffffffffu, qs, 2);
            if (g < ng) {
                uint32_t w0, w1;
                if (XSPLIT) {
                    w0 = (q[0] & 0xff) | ((q[2] & 0xff) << 8) | ((q[4] & 0xff) << 16) | ((uint32_t)(q[6] & 0xff) << 24);
                    w1 = (q[1] & 0xff) | ((q[3] & 0xff) << 8) | ((q[5] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
                } else {
                    w0 = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
                    w1 = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
                }
                const int b = g >> 2, j = g & 3;   // block, group-in-block (elements 8j..8j+7 = words 2j,2j+1)
                uint2 *dst = reinterpret_cast<uint2 *>((j < 2 ? xlo : xhi) + (size_t)b * 16 + (j & 1) * 8);
                *dst = make_uint2(w0, w1);
                if (j == 0) {
                    const float sh = round_h(scale);  // the GEMV reads the fp16 scale back (gemv.h:1631)
                    xs[b] = make_float2(sh, sh * (float)qs);
                }
            }
        }
    }
}

// Stage x as fp16 (optionally RMS-normalised) for the fp16-activation paths.
__device__ __forceinline__ void stage_x_f16(const GemvArgs &a, __half *xh, float *red, int tid) {
    const bool norm = a.norm_w != nullptr;
    float rs = 1.f;
    if (norm) rs = rms_scale_256(a.x, a.cx, a.eps, red, tid);
    const uint4 *x4 = reinterpret_cast<const uint4 *>(a.x);
    const uint4 *w4 = reinterpret_cast<const uint4 *>(a.norm_w);
    uint4 *o4 = reinterpret_cast<uint4 *>(xh);
    const int ng = a.cx >> 3;
#pragma unroll 4
    for (int g = tid; g < ng; g += GEMV_CWARPS * 32) {
        uint4 xv = x4[g];
        if (norm) {
            const uint4 wv = w4[g];
            __half2 *h = reinterpret_cast<__half2 *>(&xv);
            const __half2 *hw = reinterpret_cast<const __half2 *>(&wv);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float2 f = __half22float2(h[i]);
                const float2 fw = __half22float2(hw[i]);
                f.x = f.x * rs * (a.norm_base + fw.x);
                f.y = f.y * rs * (a.norm_base + fw.y);
                h[i] = __floats2half2_rn(f.x, f.y);
            }
        }
        o4[g] = xv;
    }
}

// ---------------------------------------------------------------------------------------
// row dot products
// ---------------------------------------------------------------------------------------
// dp4a path: kernels/gemv.h:1500-1709.  per weight block b and activation block(s):
//   y += s_x * (scale_w * sum(q_w q_x) + base_w * sum(q_x))
template <class Fmt>
__device__ __forceinline__ float row_dot_q8(const uint8_t *row, int nblk, const uint4 *xlo,
                                            const uint4 *xhi, const float2 *xs, int lane) {
    float acc = 0.f;
    for (int b = lane; b < nblk; b += 32) {
        uint32_t codes[Fmt::NW];
        float sc, ba;
        Fmt::decode(row + (size_t)b * Fmt::BYTES, codes, sc, ba);
#pragma unroll
        for (int h = 0; h < Fmt::CAP / 32; h++) {
            const int xb = b * (Fmt::CAP / 32) + h;
            const uint4 xa = xlo[xb], xc = xhi[xb];
            const float2 s = xs[xb];
            int isum = 0;
            if (Fmt::SIGNED) {
                isum = dp4a_ss(codes[8 * h + 0], xa.x, isum);
                isum = dp4a_ss(codes[8 * h + 1], xa.y, isum);
                isum = dp4a_ss(codes[8 * h + 2], xa.z, isum);
                isum = dp4a_ss(codes[8 * h + 3], xa.w, isum);
                isum = dp4a_ss(codes[8 * h + 4], xc.x, isum);
                isum = dp4a_ss(codes[8 * h + 5], xc.y, isum);
                isum = dp4a_ss(codes[8 * h + 6], xc.z, isum);
                isum = dp4a_ss(codes[8 * h + 7], xc.w, isum);
            } else {
                isum = dp4a_us(codes[8 * h + 0], xa.x, isum);
                isum = dp4a_us(codes[8 * h + 1], xa.y, isum);
                isum = dp4a_us(codes[8 * h + 2], xa.z, isum);
                isum = dp4a_us(codes[8 * h + 3], xa.w, isum);
                isum = dp4a_us(codes[8 * h + 4], xc.x, isum);
                isum = dp4a_us(codes[8 * h + 5], xc.y, isum);
                isum = dp4a_us(codes[8 * h + 6], xc.z, isum);
                isum = dp4a_us(codes[8 * h + 7], xc.w, isum);
            }
            acc = fmaf(sc * (float)isum, s.x, acc);
            if (Fmt::HAS_BASE) acc = fmaf(ba, s.y, acc);
        }
    }
    return warp_sum(acc);
}

// fp16-activation path for any block format (kernels/gemv.h:633-1470): the reference
// dequantizes a block to half and accumulates half*half products in fp32.
__device__ __forceinline__ float row_dot_hx(int dt, const uint8_t *row, int nblk, int cap, int bb,
                                            const __half *xh, int lane) {
    float acc = 0.f;
    for (int b = lane; b < nblk; b += 32) {
        const __half *xb = xh + (size_t)b * cap;
        dequant_block(dt, row + (size_t)b * bb, [&](int i, float v) {
            acc = fmaf(round_h(v), __half2float(xb[i]), acc);
        });
    }
    return warp_sum(acc);
}

// fp16 weights (GemvHalf_AX_Alg3_Kernel, gemv.h:470; used for lm_head when L > 20).
// fp32 accumulation (the reference accumulates in half -- strictly less accurate).
__device__ __forceinline__ float row_dot_f16(const uint8_t *row, int cx, const __half *xh,
                                             int lane) {
    const uint4 *w4 = reinterpret_cast<const uint4 *>(row);
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xh);
    float acc = 0.f;
    const int n8 = cx >> 3;
    for (int i = lane; i < n8; i += 32) {
        uint4 wv = w4[i], xv = x4[i];
        const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
        const __half2 *xx = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float2 wf = __half22float2(wh[k]), xf = __half22float2(xx[k]);
            acc = fmaf(wf.x, xf.x, acc);
            acc = fmaf(wf.y, xf.y, acc);
        }
    }
    for (int i = (n8 << 3) + lane; i < cx; i += 32)  // tail (cx % 8)
        acc = fmaf(__half2float(reinterpret_cast<const __half *>(row)[i]), __half2float(xh[i]), acc);
    return warp_sum(acc);
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
enum { MODE_Q8 = 0, MODE_HX = 1, MODE_F16 = 2 };

template <class Fmt, int MODE, bool GLU>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_stream_kernel(const GemvArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[GEMV_MAX_SLOTS], empty_bar[GEMV_MAX_SLOTS];
    __shared__ float red[GEMV_CWARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *ring = smem;
    uint8_t *xbase = smem + (size_t)a.nslots * a.slot_stride;
    const int nxb = a.cx >> 5;
    // MODE_Q8: xlo | xhi | xs ; else: xh (fp16)
    uint8_t *xlo = xbase, *xhi = xbase + (size_t)nxb * 16;
    float2 *xs = reinterpret_cast<float2 *>(xbase + (size_t)nxb * 32);
    __half *xh = reinterpret_cast<__half *>(xbase);

    if (tid == 0) {
        for (int i = 0; i < a.nslots; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], a.wpg);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // PDL: let the next kernel's CTAs become resident as early as possible -- their producer
    // warps only touch weights, everything else sits behind their own griddepcontrol.wait.
    pdl_launch_dependents();

    const int nmat = GLU ? 2 : 1;
    const uint32_t mat_bytes = (uint32_t)a.R * (uint32_t)a.row_bytes;

    if (warp == GEMV_CWARPS) {
        // ===== producer: stream this CTA's units; independent of the previous kernel =====
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            int j = 0;
            for (int u = blockIdx.x; u < a.units; u += gridDim.x, j++) {
                const int pos = j % a.nslots;
                mbar_wait(&empty_bar[pos], ((j / a.nslots) & 1) ^ 1);
                int seg = 0;
                if (a.nseg > 1 && u >= a.unit0[1]) seg = 1;
                if (a.nseg > 2 && u >= a.unit0[2]) seg = 2;
                const size_t off = (size_t)(u - a.unit0[seg]) * mat_bytes;
                uint8_t *dst = ring + (size_t)pos * a.slot_stride;
                mbar_arrive_expect_tx(&full_bar[pos], mat_bytes * nmat);
                bulk_g2s_hint(dst, a.w[seg] + off, mat_bytes, &full_bar[pos], pol);
                if (GLU) bulk_g2s_hint(dst + mat_bytes, a.w_gate + off, mat_bytes, &full_bar[pos], pol);
            }
        }
        return;
    }

    // ===== consumers =====
    pdl_wait();  // x / residual come from the previous kernel
    if (MODE == MODE_Q8)
        stage_x_q8<Fmt::XSPLIT>(a, xlo, xhi, xs, red, tid);
    else
        stage_x_f16(a, xh, red, tid);
    named_bar_sync(BAR_CONSUMERS, GEMV_CWARPS * 32);

    const int groups = GEMV_CWARPS / a.wpg;
    const int grp = warp / a.wpg, wl = warp % a.wpg;
    const int cap = block_capacity(a.w_dtype), bb = block_bytes(a.w_dtype);

    int j = 0;
    for (int u = blockIdx.x; u < a.units; u += gridDim.x, j++) {
        if (j % groups != grp) continue;
        const int pos = j % a.nslots;
        mbar_wait(&full_bar[pos], (j / a.nslots) & 1);
        const uint8_t *slot = ring + (size_t)pos * a.slot_stride;
        int seg = 0;
        if (a.nseg > 1 && u >= a.unit0[1]) seg = 1;
        if (a.nseg > 2 && u >= a.unit0[2]) seg = 2;
        int out_base = (u - a.unit0[seg]) * a.R;
        for (int s = 0; s < seg; s++) out_base += a.rows[s];

        for (int r = wl; r < a.R; r += a.wpg) {
            const uint8_t *row = slot + (size_t)r * a.row_bytes;
            float acc, acc2 = 0.f;
            // lane 0 fetches the epilogue operands while the row is being reduced
            float res_v = 0.f, bias_v = 0.f;
            if (lane == 0) {
                if (a.residual != nullptr) res_v = __half2float(a.residual[out_base + r]);
                if (a.bias != nullptr) bias_v = __half2float(a.bias[out_base + r]);
            }
            if (MODE == MODE_Q8) {
                acc = row_dot_q8<Fmt>(row, a.nblk, (const uint4 *)xlo, (const uint4 *)xhi, xs, lane);
                if (GLU) acc2 = row_dot_q8<Fmt>(row + mat_bytes, a.nblk, (const uint4 *)xlo,
                                                (const uint4 *)xhi, xs, lane);
            } else if (MODE == MODE_HX) {
                acc = row_dot_hx(a.w_dtype, row, a.nblk, cap, bb, xh, lane);
                if (GLU) acc2 = row_dot_hx(a.w_dtype, row + mat_bytes, a.nblk, cap, bb, xh, lane);
            } else {
                acc = row_dot_f16(row, a.cx, xh, lane);
                if (GLU) acc2 = row_dot_f16(row + mat_bytes, a.cx, xh, lane);
            }
            if (lane == 0) {
                const int o = out_base + r;
                float v = round_h(acc);  // the GEMV output tensor is half
                if (a.bias != nullptr) v = round_h(v + bias_v);
                if (GLU) {
                    // SiluActivation_Kernel (unary_tensor_opr.h:553) then Mul, each stored as half
                    float g = round_h(v / (1.0f + expf(-v)));
                    v = round_h(g * round_h(acc2));
                }
                if (a.residual != nullptr) v = round_h(v + res_v);
                a.y[o] = __float2half_rn(v);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pos]);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <class Fmt, int MODE, bool GLU>
static int launch_variant(const GemvArgs &a, size_t smem, int grid, bool pdl, cudaStream_t st) {
    auto kern = gemv_stream_kernel<Fmt, MODE, GLU>;
    static bool attr_done = false;  // per instantiation
    if (!attr_done) {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMV_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    IFB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return IFB_OK;
}

template <bool GLU>
static int dispatch(const GemvArgs &a, int mode, size_t smem, int grid, bool pdl, cudaStream_t st) {
    if (mode == MODE_F16) return launch_variant<FmtQ4_B32T1, MODE_F16, GLU>(a, smem, grid, pdl, st);
    if (mode == MODE_HX) return launch_variant<FmtQ4_B32T1, MODE_HX, GLU>(a, smem, grid, pdl, st);
    switch (a.w_dtype) {
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: return launch_variant<FmtQ4_B32T1, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        case DT_Q4_B64T1: return launch_variant<FmtQ4_B64T1, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        case DT_Q8_B32T2: return launch_variant<FmtQ8_B32T2, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        case DT_Q6_B64T1: return launch_variant<FmtQ6_B64T1, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        case DT_Q5_B64T1: return launch_variant<FmtQ5_B64T1, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        case DT_Q3H_B64T1: return launch_variant<FmtQ3H_B64T1, MODE_Q8, GLU>(a, smem, grid, pdl, st);
        default:
            set_last_error("gemv: weight type %d has no dp4a kernel (reference: Gemv_AX_QuantQ8, "
                           "tensor_mul.cu:714)", a.w_dtype);
            return IFB_ERR_TYPE;
    }
}

static bool is_ax8_type(int dt) {
    switch (dt) {
        case DT_Q8_B32T2: case DT_Q6_B64T1: case DT_Q5_B64T1: case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: case DT_Q4_B64T1: case DT_Q3H_B64T1: return true;
        default: return false;
    }
}

int gemv_launch(const GemvLaunch &g, cudaStream_t st) {
    const int dt = g.w_dtype;
    const int cap = block_capacity(dt), bb = block_bytes(dt);
    IFB_CHECK_ARG(cap > 0 && dt != DT_F32, "gemv: unsupported weight type %d", dt);
    IFB_CHECK_ARG(g.nseg >= 1 && g.nseg <= 3, "gemv: nseg %d", g.nseg);
    IFB_CHECK_ARG(g.cx > 0 && g.cx % 32 == 0,
                  "gemv: cx (%lld) should be a multiple of 32", (long long)g.cx);
    if (dt != DT_F16) IFB_CHECK_ARG(g.cx % cap == 0, "gemv: cx %% block capacity != 0");
    IFB_CHECK_ARG(!(g.w_gate != nullptr && g.nseg != 1), "gemv: GLU needs exactly one segment");
    IFB_CHECK_ARG((((uintptr_t)g.x | (uintptr_t)g.norm_w) & 15) == 0,
                  "gemv: x / norm weight must be 16-byte aligned");

    int mode;
    if (dt == DT_F16) {
        mode = MODE_F16;
        IFB_CHECK_ARG(g.xq == nullptr, "gemv: fp16 weights need an fp16 activation");
    } else if ((g.quantize_x || g.xq != nullptr)) {
        if (!is_ax8_type(dt)) {
            set_last_error("gemv: element type %d is not supported on the Q8-activation path "
                           "(reference: Gemv_AX_QuantQ8)", dt);
            return IFB_ERR_TYPE;
        }
        mode = MODE_Q8;
    } else {
        mode = MODE_HX;
    }

    GemvArgs a = {};
    a.w_dtype = dt;
    a.cx = (int)g.cx;
    a.nblk = (dt == DT_F16) ? (int)(g.cx / 32) : (int)(g.cx / cap);
    a.row_bytes = (int)row_bytes(dt, g.cx);
    a.nseg = g.nseg;
    a.w_gate = (const uint8_t *)g.w_gate;
    a.x = (const __half *)g.x;
    a.xq = (const uint8_t *)g.xq;
    a.norm_w = (const __half *)g.norm_w;
    a.eps = g.norm_eps;
    a.norm_base = g.norm_base;
    a.residual = (const __half *)g.residual;
    a.bias = (const __half *)g.bias;
    a.y = (__half *)g.y;

    // rows per unit: R*row_bytes must be a multiple of 16 (bulk-copy granularity) and divide
    // every segment; prefer 8 (one row per consumer warp), fall back to 4 when a slot would
    // exceed ~40 KB, go to 16/32 for very short rows.
    const int nmat = g.w_gate ? 2 : 1;
    int gcd_rows = 0;
    for (int i = 0; i < g.nseg; i++) {
        IFB_CHECK_ARG(g.rows[i] > 0 && g.w[i] != nullptr, "gemv: empty segment %d", i);
        IFB_CHECK_ARG(((uintptr_t)g.w[i] & 15) == 0, "gemv: weight pointer not 16-byte aligned");
        int r = g.rows[i];
        int x = gcd_rows, y = r;
        while (y) { int t = x % y; x = y; y = t; }
        gcd_rows = gcd_rows == 0 ? r : x;
    }
    auto valid = [&](int c) { return gcd_rows % c == 0 && ((int64_t)c * a.row_bytes) % 16 == 0; };
    auto sbytes = [&](int c) { return (int64_t)c * a.row_bytes * nmat; };
    int R = 0;
    if (valid(8) && sbytes(8) <= 40 * 1024) R = 8;
    else if (valid(4)) R = 4;
    else if (valid(8)) R = 8;
    if (R == 8 && valid(16) && sbytes(16) <= 16 * 1024) R = 16;   // very short rows: fatter copies
    if (R == 16 && valid(32) && sbytes(32) <= 16 * 1024) R = 32;
    IFB_CHECK_ARG(R != 0, "gemv: rows (gcd %d) x row_bytes %d cannot be tiled in 16-byte units "
                  "(rows must be a multiple of 4, or 8 for 34-byte blocks)", gcd_rows, a.row_bytes);
    a.R = R;
    a.wpg = R < GEMV_CWARPS ? R : GEMV_CWARPS;

    int units = 0;
    for (int i = 0; i < g.nseg; i++) {
        a.w[i] = (const uint8_t *)g.w[i];
        a.rows[i] = g.rows[i];
        a.unit0[i] = units;
        units += g.rows[i] / R;
    }
    a.units = units;

    const int64_t slot_bytes = (int64_t)R * a.row_bytes * nmat;
    a.slot_stride = (int)((slot_bytes + 16 + 127) / 128 * 128);
    const int64_t xbytes = (mode == MODE_Q8) ? (g.cx / 32) * 40 : g.cx * 2;
    const int64_t xpad = (xbytes + 127) / 128 * 128;
    int64_t budget = (g.smem_budget > 0 ? g.smem_budget : 100 * 1024);
    int nslots = (int)((budget - xpad) / a.slot_stride);
    if (nslots < 2) {
        budget = 216 * 1024;
        nslots = (int)((budget - xpad) / a.slot_stride);
    }
    IFB_CHECK_ARG(nslots >= 1, "gemv: a %lld-byte slot does not fit shared memory", (long long)slot_bytes);
    if (nslots > GEMV_MAX_SLOTS) nslots = GEMV_MAX_SLOTS;
    a.nslots = nslots;
    const size_t smem = (size_t)nslots * a.slot_stride + xpad;

    const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    int grid = sm_count() * ctas_per_sm;
    if (grid > units) grid = units;

    if (g.w_gate) return dispatch<true>(a, mode, smem, grid, g.pdl, st);
    return dispatch<false>(a, mode, smem, grid, g.pdl, st);
}

}  // namespace ifb
