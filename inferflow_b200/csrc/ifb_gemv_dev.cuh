// Device functions of the dequant-fused GEMV shared by the stand-alone streaming kernel (ifb_gemv.cu) and the
// persistent decode-step kernel (ifb_step.cu): activation staging (RMSNorm + Q8_B32T2 quantization with the
// reference's arithmetic, kernels/tensor_quant.h:45-82, unary_tensor_opr.h:217) and the per-row dot products
// (kernels/gemv.h:470-1709). One copy of the arithmetic: both kernels produce the same bits.
#pragma once
#include "ifb_formats.cuh"

namespace ifb {

constexpr int GEMV_CWARPS = 16;                       // consumer warps
constexpr int GEMV_CTHREADS = GEMV_CWARPS * 32;
constexpr int BAR_CONSUMERS = 1;                      // named barrier id for the consumer warps

// what the activation staging needs to know
struct XStage {
    int cx;
    const uint8_t *xq;        // pre-quantized activation (Q8_B32T2 blocks) or null
    const __half *norm_w;     // RMSNorm weight or null
    float eps, norm_base;
};

// 128-bit load of staging input. LD = 0: ordinary load (kernel-boundary coherence: stand-alone kernels; shared
// memory sources); LD = 1: ld.global.cg -- L2 only, for vectors another CTA of the SAME persistent kernel wrote
// (the L1 is not invalidated inside a launch).
template <int LD>
__device__ __forceinline__ uint4 ld_x128(const uint4 *p) {
    if (LD == 0) return *p;
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

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
// slot of activation block b in the staged arrays. VEC4 kernels read, per pass of 128 blocks,
// block 4*lane+i with all lanes at once: store those 32 blocks contiguously (transpose 4x32).
template <bool VEC4>
__device__ __forceinline__ int x_slot(int b) {
    if (!VEC4) return b;
    return (b & ~127) | ((b & 3) << 5) | ((b & 127) >> 2);
}

// RMS scale of the whole vector, computed by the consumer threads (128-bit loads).
// Tensor_RmsNorm_Kernel (unary_tensor_opr.h:217): scale = 1/sqrt(mean(x^2) + eps)
template <int LD = 0, int NW = GEMV_CWARPS>
__device__ __forceinline__ float rms_scale(const __half *x, int cx, float eps, float *red, int tid) {
    constexpr int NT = NW * 32;   // consumer threads taking part (the stand-alone kernel: 16 warps; the step kernel: 15)
    const uint4 *x4 = reinterpret_cast<const uint4 *>(x);
    const int ng = cx >> 3;
    float s = 0.f;
#pragma unroll 4
    for (int g = tid; g < ng; g += NT) {
        const uint4 v = ld_x128<LD>(x4 + g);
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
    named_bar_sync(BAR_CONSUMERS, NT);
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NW; i++) t += red[i];
    const float mean = t / (float)cx;
    return 1.0f / sqrtf(mean + eps);
}

// Stage x as Q8_B32T2 (codes + fp16-rounded scale + scale*sum(q)) into shared memory.
// The arithmetic per block of 32 is the reference's alg 2 (tensor_quant.h:45-82): scale =
// absmax/127, q = roundf(v/scale), scale <= 1e-6 -> 0. Mapping: one thread = 8 consecutive
// elements (ONE 128-bit load), four neighbouring lanes = one block; all of a thread's loads are
// issued before any is used, so the prologue costs one memory round trip, not one per block.
template <bool XSPLIT, bool VEC4, int LD = 0, int NW = GEMV_CWARPS>
__device__ __forceinline__ void stage_x_q8(const XStage &a, const void *xsrc, uint8_t *xlo, uint8_t *xhi,
                                           float2 *xs, float *red, int tid) {
    constexpr int NT = NW * 32;
    const int warp = tid >> 5, lane = tid & 31;
    const int nxb = a.cx >> 5;
    if (a.xq != nullptr) {  // already quantized by the caller (TensorMul::Gemv_AX with a Q8 X)
        for (int b = warp; b < nxb; b += NW) {
            const uint8_t *blk = a.xq + (size_t)b * 34;
            const int q = (int)(int8_t)blk[2 + lane];
            const int p = x_pos<XSPLIT>(lane), sl = x_slot<VEC4>(b);
            (p < 16 ? xlo + sl * 16 + p : xhi + sl * 16 + (p - 16))[0] = (uint8_t)q;
            const int qs = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) {
                const float s = h2f(*(const uint16_t *)blk);
                xs[sl] = make_float2(s, s * (float)qs);
            }
        }
        return;
    }
    const bool norm = a.norm_w != nullptr;
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xsrc);
    const uint4 *w4 = reinterpret_cast<const uint4 *>(a.norm_w);
    const int ng = a.cx >> 3;                 // groups of 8 elements; ng % 4 == 0
    constexpr int KPT = 3;                    // groups per thread per round
    const bool one_round = ng <= NT * KPT;   // cx <= 12288: x is read from memory exactly once
    float rs = 1.f;
    if (norm && !one_round) rs = rms_scale<LD, NW>(reinterpret_cast<const __half *>(xsrc), a.cx, a.eps, red, tid);
    for (int g0 = 0; g0 < ng; g0 += NT * KPT) {
        uint4 xv[KPT], wv[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * NT + tid;
            xv[k] = make_uint4(0, 0, 0, 0);
            wv[k] = make_uint4(0, 0, 0, 0);
            if (g < ng) {
                xv[k] = ld_x128<LD>(x4 + g);
                if (norm) wv[k] = w4[g];
            }
        }
        if (norm && one_round) {
            // Tensor_RmsNorm_Kernel (unary_tensor_opr.h:217) from the registers just loaded
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const __half2 *h = reinterpret_cast<const __half2 *>(&xv[k]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __half22float2(h[i]);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) red[warp] = ss;
            named_bar_sync(BAR_CONSUMERS, NT);
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < NW; i++) t += red[i];
            rs = 1.0f / sqrtf(t / (float)a.cx + a.eps);
        }
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * NT + tid;   // a quad of lanes shares one block
            if (g0 + k * NT >= ng) break;      // uniform: the whole sub-round is past the end
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
            const bool dead = scale <= 0.000001f;
            const float inv = dead ? 0.f : __frcp_rn(scale);
            int q[8], qs = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int t = dead ? 0 : q8_round_div(v[i], scale, inv);
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
                const int sl = x_slot<VEC4>(b);
                uint2 *dst = reinterpret_cast<uint2 *>((j < 2 ? xlo : xhi) + (size_t)sl * 16 + (j & 1) * 8);
                *dst = make_uint2(w0, w1);
                if (j == 0) {
                    const float sh = round_h(scale);  // the GEMV reads the fp16 scale back (gemv.h:1631)
                    xs[sl] = make_float2(sh, sh * (float)qs);
                }
            }
        }
    }
}

// The same staging, small: ONE instantiation serves every layout / source (vec4 and ldcg are run-time flags) and the
// per-group quantization is a rolled loop. For the persistent step kernel, where this code runs once per op boundary and
// is cold every time (the unrolled form above is ~30 KB of SASS per instantiation; the instruction fetch, not the
// arithmetic, was what a boundary cost). Bit-for-bit the arithmetic of stage_x_q8.
template <bool XSPLIT, int NW>
__device__ __forceinline__ void stage_x_q8_compact(const XStage &a, const void *xsrc, uint8_t *xlo, uint8_t *xhi,
                                                   float2 *xs, float *red, int tid, bool vec4, bool ldcg) {
    constexpr int NT = NW * 32;
    const int warp = tid >> 5, lane = tid & 31;
    const bool norm = a.norm_w != nullptr;
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xsrc);
    const uint4 *w4 = reinterpret_cast<const uint4 *>(a.norm_w);
    const int ng = a.cx >> 3;
    constexpr int KPT = 3;
    float ss_total = 0.f;
    float rs = 1.f;
    const bool one_round = ng <= NT * KPT;
    if (norm && !one_round) {
        // Tensor_RmsNorm_Kernel (unary_tensor_opr.h:217): scale = 1/sqrt(mean(x^2) + eps), a pass of its own
        float sq = 0.f;
        for (int g = tid; g < ng; g += NT) {
            const uint4 v = ldcg ? ld_x128<1>(x4 + g) : x4[g];
            const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 f = __half22float2(h[k]);
                sq = fmaf(f.x, f.x, sq);
                sq = fmaf(f.y, f.y, sq);
            }
        }
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        named_bar_sync(BAR_CONSUMERS, NT);
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NW; i++) t += red[i];
        rs = 1.0f / sqrtf(t / (float)a.cx + a.eps);
    }
    (void)ss_total;
    for (int g0 = 0; g0 < ng; g0 += NT * KPT) {
        uint4 xv0 = make_uint4(0, 0, 0, 0), xv1 = xv0, xv2 = xv0, wv0 = xv0, wv1 = xv0, wv2 = xv0;
        {   // all loads of the round first: one memory round trip
            const int ga = g0 + tid, gb = ga + NT, gc = gb + NT;
            if (ga < ng) { xv0 = ldcg ? ld_x128<1>(x4 + ga) : x4[ga]; if (norm) wv0 = w4[ga]; }
            if (gb < ng) { xv1 = ldcg ? ld_x128<1>(x4 + gb) : x4[gb]; if (norm) wv1 = w4[gb]; }
            if (gc < ng) { xv2 = ldcg ? ld_x128<1>(x4 + gc) : x4[gc]; if (norm) wv2 = w4[gc]; }
        }
        if (norm && one_round) {
            float ss = 0.f;
            const uint4 xs3[3] = {xv0, xv1, xv2};
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const __half2 *h = reinterpret_cast<const __half2 *>(&xs3[k]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __half22float2(h[i]);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) red[warp] = ss;
            named_bar_sync(BAR_CONSUMERS, NT);
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < NW; i++) t += red[i];
            rs = 1.0f / sqrtf(t / (float)a.cx + a.eps);
        }
#pragma unroll 1
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * NT + tid;   // a quad of lanes shares one block
            if (g0 + k * NT + warp * 32 >= ng) break;   // warp-uniform (ng % 32 == 0 for cx % 256 == 0; else the tail warp
                                                        // runs with idle lanes): nothing left for this warp
            const uint4 xk = k == 0 ? xv0 : (k == 1 ? xv1 : xv2);
            const uint4 wk = k == 0 ? wv0 : (k == 1 ? wv1 : wv2);
            float v[8];
            {
                const __half2 *h = reinterpret_cast<const __half2 *>(&xk);
                const __half2 *hw = reinterpret_cast<const __half2 *>(&wk);
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
            const bool dead = scale <= 0.000001f;
            const float inv = dead ? 0.f : __frcp_rn(scale);
            int q[8], qs = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int t = dead ? 0 : q8_round_div(v[i], scale, inv);
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
                const int sl = vec4 ? x_slot<true>(b) : b;
                uint2 *dst = reinterpret_cast<uint2 *>((j < 2 ? xlo : xhi) + (size_t)sl * 16 + (j & 1) * 8);
                *dst = make_uint2(w0, w1);
                if (j == 0) {
                    const float sh = round_h(scale);  // the GEMV reads the fp16 scale back (gemv.h:1631)
                    xs[sl] = make_float2(sh, sh * (float)qs);
                }
            }
        }
    }
}

// Stage x as fp16 (optionally RMS-normalised) for the fp16-activation paths.
template <int LD = 0, int NW = GEMV_CWARPS>
__device__ __forceinline__ void stage_x_f16(const XStage &a, const void *xsrc, __half *xh, float *red, int tid) {
    constexpr int NT = NW * 32;
    const bool norm = a.norm_w != nullptr;
    float rs = 1.f;
    if (norm) rs = rms_scale<LD, NW>(reinterpret_cast<const __half *>(xsrc), a.cx, a.eps, red, tid);
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xsrc);
    const uint4 *w4 = reinterpret_cast<const uint4 *>(a.norm_w);
    uint4 *o4 = reinterpret_cast<uint4 *>(xh);
    const int ng = a.cx >> 3;
#pragma unroll 4
    for (int g = tid; g < ng; g += NT) {
        uint4 xv = ld_x128<LD>(x4 + g);
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
// row dot products (per-lane partial sums; the caller reduces across the warp)
// ---------------------------------------------------------------------------------------
// dp4a path: kernels/gemv.h:1500-1709.  per weight block b and activation block(s):
//   y += s_x * (scale_w * sum(q_w q_x) + base_w * sum(q_x))
template <class Fmt>
__device__ __forceinline__ float row_dot_q8(const uint8_t *row, int nblk, const uint4 *xlo,
                                            const uint4 *xhi, const float2 *xs, int lane, float acc = 0.f) {
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
    return acc;
}

// One Q4_B32T1 block held in registers (hdr = base|scale<<16, d0..d3 = 32 nibbles) against one
// staged activation block (nibble-split order). Low nibbles: q ; high nibbles stay in place as
// 16*q (<= 240 fits a u8 lane of dp4a) and are scaled back with one arithmetic shift.
__device__ __forceinline__ float q4b32_block_dot(uint32_t hdr, uint32_t d0, uint32_t d1, uint32_t d2,
                                                 uint32_t d3, const uint4 xa, const uint4 xc,
                                                 const float2 s, float acc) {
    int lo = dp4a_us(d0 & 0x0f0f0f0fu, xa.x, 0);
    int hi = dp4a_us(d0 & 0xf0f0f0f0u, xa.y, 0);
    lo = dp4a_us(d1 & 0x0f0f0f0fu, xa.z, lo);
    hi = dp4a_us(d1 & 0xf0f0f0f0u, xa.w, hi);
    lo = dp4a_us(d2 & 0x0f0f0f0fu, xc.x, lo);
    hi = dp4a_us(d2 & 0xf0f0f0f0u, xc.y, hi);
    lo = dp4a_us(d3 & 0x0f0f0f0fu, xc.z, lo);
    hi = dp4a_us(d3 & 0xf0f0f0f0u, xc.w, hi);
    const int isum = lo + (hi >> 4);
    const float2 bs = __half22float2(*reinterpret_cast<const __half2 *>(&hdr));  // x = base, y = scale
    acc = fmaf(bs.y * (float)isum, s.x, acc);
    return fmaf(bs.x, s.y, acc);
}

// Q4_B32T1 rows with nblk % 4 == 0: a lane owns blocks 4*lane..4*lane+3 of each pass of 128
// (80 contiguous bytes = 5 LDS.128, bank-conflict free across the warp). NM = 2 processes the
// gate row of the GLU pair against the same activation registers. ZERO = false continues the sums in acc (a row
// taken in several K-chunks: same per-lane order of additions as the whole row at once).
template <int NM, bool ZERO = true>
__device__ __forceinline__ void row_dot_q4b32_vec4(const uint8_t *row, size_t mat_bytes, int nblk,
                                                   const uint4 *xlo, const uint4 *xhi,
                                                   const float2 *xs, int lane, float *acc) {
    if (ZERO) {
#pragma unroll
        for (int m = 0; m < NM; m++) acc[m] = 0.f;
    }
    for (int p0 = 0; p0 < nblk; p0 += 128) {
        const int b0 = p0 + 4 * lane;
        if (b0 < nblk) {
            uint4 xa[4], xc[4];
            float2 s[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sl = p0 + i * 32 + lane;
                xa[i] = xlo[sl];
                xc[i] = xhi[sl];
                s[i] = xs[sl];
            }
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const uint4 *wp = reinterpret_cast<const uint4 *>(row + m * mat_bytes + (size_t)b0 * 20);
                const uint4 q0 = wp[0], q1 = wp[1], q2 = wp[2], q3 = wp[3], q4 = wp[4];
                float a = acc[m];
                a = q4b32_block_dot(q0.x, q0.y, q0.z, q0.w, q1.x, xa[0], xc[0], s[0], a);
                a = q4b32_block_dot(q1.y, q1.z, q1.w, q2.x, q2.y, xa[1], xc[1], s[1], a);
                a = q4b32_block_dot(q2.z, q2.w, q3.x, q3.y, q3.z, xa[2], xc[2], s[2], a);
                a = q4b32_block_dot(q3.w, q4.x, q4.y, q4.z, q4.w, xa[3], xc[3], s[3], a);
                acc[m] = a;
            }
        }
    }
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
    return acc;
}

// fp16 weights (GemvHalf_AX_Alg3_Kernel, gemv.h:470; used for lm_head when L > 20).
// fp32 accumulation (the reference accumulates in half -- strictly less accurate).
__device__ __forceinline__ float row_dot_f16(const uint8_t *row, int cx, const __half *xh,
                                             int lane, float acc = 0.f) {
    const uint4 *w4 = reinterpret_cast<const uint4 *>(row);
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xh);
    const int n8 = cx >> 3;
#pragma unroll 4
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
    return acc;
}

}  // namespace ifb
