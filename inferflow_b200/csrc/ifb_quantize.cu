// Block quantizers / dequantizers on the device.
//
// Replaces TensorOpr::Quantize / TensorOpr::Dequantize (reference src/tensor/tensor_opr.cu:1623,
// kernels src/kernels/tensor_quant.h:9-868) whose per-block arithmetic is
// Quantization::QuantizeRow_* / Dequantize* (src/common/quantization.h:110-1194).
// Codes and fp16 parameters are bit-exact with the reference AS BUILT: both of its flavours
// contract `(v-min)*inv + 0.5f` and `q*scale + base` into fused multiply-adds (nvcc --fmad=true;
// g++ -O2 -mfma), so the fused forms are written out explicitly here.
//
// Load-time work except for Q8_B32T2 (activations, KV rows), which has the reference's
// warp-per-block algorithm (Tensor_QuantizeQ8_B32T2_Alg2_Kernel, tensor_quant.h:45-82).
#include "ifb_formats.cuh"
#include "ifb_kernels.h"

namespace ifb {

struct QParams {
    float mn, scale, inv, base;
};

// GetValueRange (quantization.h:71-87) + scale = (max-min)/div, inv = scale >= 1e-5 ? 1/scale : 0
__device__ __forceinline__ QParams lin_params(const float *v, int n, float div, int mode) {
    float mn = v[0], mx = v[0];
    for (int i = 1; i < n; i++) {
        if (mn > v[i]) mn = v[i];
        if (mx < v[i]) mx = v[i];
    }
    QParams p;
    p.mn = mn;
    p.scale = __fdiv_rn(__fsub_rn(mx, mn), div);
    p.inv = p.scale >= 0.00001f ? __fdiv_rn(1.0f, p.scale) : 0.0f;
    p.base = mode ? __fmaf_rn(0.5f, p.scale, mn) : mn;
    return p;
}
// mode 0: (uint32)(fma(v-min, inv, 0.5)) ; mode 1 ("B" variants): (uint32)(fma(v-min, inv, 1e-4))
__device__ __forceinline__ uint32_t lin_q(float v, const QParams &p, int mode, uint32_t maxq) {
    float u = __fmaf_rn(__fsub_rn(v, p.mn), p.inv, mode ? 0.0001f : 0.5f);
    uint32_t q = (uint32_t)u;  // cvt.rzi.u32.f32 saturates negatives to 0 like the x86 build's range
    return q > maxq ? maxq : q;
}
// Block headers sit at even offsets: store fp16 bit patterns with ONE 16-bit store. (Splitting
// them into two byte stores makes ptxas 12.9 emit F2I.U8.F16 for the low byte -- it converts the
// half VALUE instead of truncating its bits.)
__device__ __forceinline__ void put16(uint8_t *p, uint16_t v) {
    *reinterpret_cast<uint16_t *>(p) = v;
}

// one thread = one block (load-time path, mirrors the reference's launch shape)
__global__ void quantize_blocks_kernel(int dt, uint8_t *dst, const __half *src, int64_t nblocks,
                                       int cap, int bb) {
    int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nblocks) return;
    float s[64];
    const __half *sp = src + bi * cap;
    for (int i = 0; i < cap; i++) s[i] = __half2float(sp[i]);
    uint8_t *b = dst + bi * bb;
    switch (dt) {
        case DT_Q8_B32T1: {  // quantization.h:110-154
            QParams p = lin_params(s, 32, 255.0f, 0);
            put16(b, f2h_bits(p.mn));
            put16(b + 2, f2h_bits(p.scale));
            for (int i = 0; i < 32; i++) b[4 + i] = (uint8_t)lin_q(s[i], p, 0, 255);
        } break;
        case DT_Q6_B64T1: {  // :268-318 (62 steps, clamp 63)
            QParams p = lin_params(s, 64, 62.0f, 0);
            put16(b, f2h_bits(p.mn));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 16; r++) {
                uint32_t q0 = lin_q(s[4 * r], p, 0, 63), q1 = lin_q(s[4 * r + 1], p, 0, 63),
                         q2 = lin_q(s[4 * r + 2], p, 0, 63), q3 = lin_q(s[4 * r + 3], p, 0, 63);
                b[4 + r] = (uint8_t)((q0 >> 4) | ((q1 >> 4) << 2) | ((q2 >> 4) << 4) | ((q3 >> 4) << 6));
                b[20 + 2 * r] = (uint8_t)((q0 & 15) | ((q1 & 15) << 4));
                b[20 + 2 * r + 1] = (uint8_t)((q2 & 15) | ((q3 & 15) << 4));
            }
        } break;
        case DT_Q5_B64T1: {  // :446-500 (30 steps, clamp 31)
            QParams p = lin_params(s, 64, 30.0f, 0);
            put16(b, f2h_bits(p.mn));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 8; r++) {
                uint32_t q[8], h = 0;
                for (int k = 0; k < 8; k++) {
                    q[k] = lin_q(s[8 * r + k], p, 0, 31);
                    h |= ((q[k] >> 4) & 1) << k;
                }
                b[4 + r] = (uint8_t)h;
                for (int k = 0; k < 4; k++)
                    b[12 + 4 * r + k] = (uint8_t)((q[2 * k] & 15) | ((q[2 * k + 1] & 15) << 4));
            }
        } break;
        case DT_Q5_B32T1: {  // :348-395 (scale first, /31, no clamp)
            QParams p = lin_params(s, 32, 31.0f, 0);
            put16(b, f2h_bits(p.scale));
            put16(b + 2, f2h_bits(p.mn));
            uint32_t qh = 0;
            for (int r = 0; r < 16; r++) {
                uint32_t q1 = lin_q(s[r], p, 0, 0xffffffffu), q2 = lin_q(s[r + 16], p, 0, 0xffffffffu);
                b[8 + r] = (uint8_t)((q1 & 15) | ((q2 & 15) << 4));
                qh |= ((q1 & 16) >> 4) << r;
                qh |= ((q2 & 16) >> 4) << (r + 16);
            }
            b[4] = (uint8_t)qh; b[5] = (uint8_t)(qh >> 8); b[6] = (uint8_t)(qh >> 16); b[7] = (uint8_t)(qh >> 24);
        } break;
        case DT_Q4_B16: {  // :43-68, :657-708
            float mn = s[0], mx = s[0];
            for (int i = 1; i < 16; i++) { if (mn > s[i]) mn = s[i]; if (mx < s[i]) mx = s[i]; }
            // float->uint8 of an out-of-range value is UB in C++; the reference as built (x86 cvttsd2si, low byte)
            // wraps, CUDA's cvt saturates: go through int to reproduce the reference's bytes.
            uint8_t u8 = (uint8_t)(int)((double)__fmul_rn(mn, 100.0f) + 100.01);
            mn = __fsub_rn(__fdiv_rn((float)(int)(uint32_t)u8, 100.0f), 1.0f);
            QParams p;
            p.mn = mn;
            p.scale = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
            p.inv = p.scale >= 0.00001f ? __fdiv_rn(1.0f, p.scale) : 0.0f;
            p.base = mn;
            b[0] = (uint8_t)(int)__fmaf_rn(mn, 100.0f, 100.5f);
            b[1] = (uint8_t)(int)__fmaf_rn(p.scale, 1000.0f, 0.5f);
            for (int r = 0; r < 8; r++)
                b[2 + r] = (uint8_t)(lin_q(s[2 * r], p, 0, 15) | (lin_q(s[2 * r + 1], p, 0, 15) << 4));
        } break;
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: {  // A :535-587 (/15 round), B :589-631 (/16, base=min+scale/2, trunc)
            int mode = dt == DT_Q4_B32T1B;
            QParams p = lin_params(s, 32, mode ? 16.0f : 15.0f, mode);
            put16(b, f2h_bits(p.base));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 16; r++)
                b[4 + r] = (uint8_t)(lin_q(s[2 * r], p, mode, 15) | (lin_q(s[2 * r + 1], p, mode, 15) << 4));
        } break;
        case DT_Q4_B64T1: {  // :757-803 (14 steps, clamp 15)
            QParams p = lin_params(s, 64, 14.0f, 0);
            put16(b, f2h_bits(p.mn));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 32; r++)
                b[4 + r] = (uint8_t)(lin_q(s[2 * r], p, 0, 15) | (lin_q(s[2 * r + 1], p, 0, 15) << 4));
        } break;
        case DT_Q3H_B64T1: {  // :854-926 (11 levels, pairs in base 11)
            QParams p = lin_params(s, 64, 10.0f, 0);
            put16(b, f2h_bits(p.mn));
            put16(b + 2, f2h_bits(p.scale));
            b[4] = b[5] = b[6] = b[7] = 0;
            for (int r = 0; r < 8; r++) {
                int P[4];
                for (int k = 0; k < 4; k++) {
                    int q0 = (int)__fmaf_rn(__fsub_rn(s[8 * r + 2 * k], p.mn), p.inv, 0.5f);
                    int q1 = (int)__fmaf_rn(__fsub_rn(s[8 * r + 2 * k + 1], p.mn), p.inv, 0.5f);
                    q0 = min(max(q0, 0), 10);
                    q1 = min(max(q1, 0), 10);
                    P[k] = q0 + 11 * q1;
                }
                b[16 + 2 * r] = (uint8_t)((P[0] & 15) | ((P[1] & 15) << 4));
                b[16 + 2 * r + 1] = (uint8_t)((P[2] & 15) | ((P[3] & 15) << 4));
                b[8 + r] = (uint8_t)(((P[0] >> 4) & 3) | (((P[1] >> 4) & 3) << 2) |
                                     (((P[2] >> 4) & 3) << 4) | (((P[3] >> 4) & 3) << 6));
                int hn = ((P[0] >> 6) & 1) | (((P[1] >> 6) & 1) << 1) | (((P[2] >> 6) & 1) << 2) |
                         (((P[3] >> 6) & 1) << 3);
                b[4 + r / 2] |= (uint8_t)(hn << ((r & 1) * 4));
            }
        } break;
        case DT_Q3_B32T1A:
        case DT_Q3_B32T1B: {  // A :964-1014 (/7), B :1016-1068 (/8)
            int mode = dt == DT_Q3_B32T1B;
            QParams p = lin_params(s, 32, mode ? 8.0f : 7.0f, mode);
            put16(b, f2h_bits(p.base));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 4; r++) {
                uint32_t q[8], h = 0, lo = 0, hi = 0;
                for (int k = 0; k < 8; k++) {
                    q[k] = lin_q(s[8 * r + k], p, mode, 7);
                    h |= ((q[k] >> 2) & 1) << k;
                }
                for (int k = 0; k < 4; k++) {
                    lo |= (q[k] & 3) << (2 * k);
                    hi |= (q[4 + k] & 3) << (2 * k);
                }
                b[8 + 2 * r] = (uint8_t)lo;
                b[8 + 2 * r + 1] = (uint8_t)hi;
                b[4 + r] = (uint8_t)h;
            }
        } break;
        case DT_Q2_B32T1A:
        case DT_Q2_B32T1B: {  // A :1098-1146 (/3), B :1148-1194 (/4)
            int mode = dt == DT_Q2_B32T1B;
            QParams p = lin_params(s, 32, mode ? 4.0f : 3.0f, mode);
            put16(b, f2h_bits(p.base));
            put16(b + 2, f2h_bits(p.scale));
            for (int r = 0; r < 8; r++) {
                uint32_t v = 0;
                for (int k = 0; k < 4; k++) v |= lin_q(s[4 * r + k], p, mode, 3) << (2 * k);
                b[4 + r] = (uint8_t)v;
            }
        } break;
        default: break;
    }
}

// Q8_B32T2, the reference's alg 2: one warp per block, lane = element
__global__ void quantize_q8t2_kernel(uint8_t *dst, const __half *src, int64_t nblocks) {
    int64_t bi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (bi >= nblocks) return;
    float v = __half2float(src[bi * 32 + lane]);
    float m = warp_max(fabsf(v));
    float scale = __fdiv_rn(m, 127.0f);
    int q = scale <= 0.000001f ? 0 : (int)roundf(__fdiv_rn(v, scale));
    q = min(max(q, -128), 127);
    uint8_t *b = dst + bi * 34;
    b[2 + lane] = (uint8_t)(int8_t)q;
    if (lane == 0) put16(b, f2h_bits(scale));
}

__global__ void dequantize_blocks_kernel(int dt, __half *dst, const uint8_t *src, int64_t nblocks,
                                         int cap, int bb) {
    int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nblocks) return;
    __half *o = dst + bi * cap;
    dequant_block(dt, src + bi * bb, [&](int i, float v) { o[i] = __float2half_rn(v); });
}

// Q3H_B64T1 -> Q4_B64T1, block by block, LOSSLESS: the same fp16 (base, scale), the 64 codes (0..10, two per 7-bit
// pair P = q0 + 11*q1, Quantization::Dequantize of the reference's q3h blocks) rewritten as plain nibbles (element 2j
// in the low, 2j+1 in the high nibble of data byte j). value = q*scale + base is the same expression for both block
// types, so every kernel computes the same bits from either form; the nibble form costs 36 instead of 32 bytes per 64
// weights and an order of magnitude fewer instructions to decode (the pair split is a divide by 11 per two weights).
__global__ void repack_q3h_q4b64_kernel(uint8_t *dst, const uint8_t *src, int64_t nblocks) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src + b * 32);
    uint32_t *o = reinterpret_cast<uint32_t *>(dst + b * 36);
    uint32_t codes[16];
    float sc, ba;
    FmtQ3H_B64T1::decode(reinterpret_cast<const uint8_t *>(w), codes, sc, ba);   // codes[k]: bytes = elements 4k..4k+3
    o[0] = w[0];
#pragma unroll
    for (int i = 0; i < 8; i++) {       // output word i = elements 8i..8i+7, byte j = (e[2j], e[2j+1])
        const uint32_t a = codes[2 * i], c = codes[2 * i + 1];
        const uint32_t b0 = (a & 0xfu) | ((a >> 4) & 0xf0u);
        const uint32_t b1 = ((a >> 16) & 0xfu) | ((a >> 20) & 0xf0u);
        const uint32_t b2 = (c & 0xfu) | ((c >> 4) & 0xf0u);
        const uint32_t b3 = ((c >> 16) & 0xfu) | ((c >> 20) & 0xf0u);
        o[1 + i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
}

int repack_q3h_q4b64(void *dst, const void *src, int64_t nblocks, cudaStream_t st) {
    if (nblocks <= 0) return IFB_OK;
    repack_q3h_q4b64_kernel<<<(unsigned)((nblocks + 127) / 128), 128, 0, st>>>((uint8_t *)dst, (const uint8_t *)src, nblocks);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

int quantize_rows(int dt, void *dst, const void *src, int64_t rows, int64_t cx, cudaStream_t st) {
    int cap = block_capacity(dt), bb = block_bytes(dt);
    IFB_CHECK_ARG(is_quantized(dt) && cap > 1, "Quantize: element type %d is not supported", dt);
    IFB_CHECK_ARG(cx % cap == 0, "The number of columns should be a multiple of %d", cap);
    if (rows <= 0) return IFB_OK;
    int64_t nblocks = rows * (cx / cap);
    if (dt == DT_Q8_B32T2) {
        int64_t threads = nblocks * 32;
        quantize_q8t2_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
            (uint8_t *)dst, (const __half *)src, nblocks);
    } else {
        quantize_blocks_kernel<<<(unsigned)((nblocks + 127) / 128), 128, 0, st>>>(
            dt, (uint8_t *)dst, (const __half *)src, nblocks, cap, bb);
    }
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

int dequantize_rows(int dt, void *dst, const void *src, int64_t rows, int64_t cx, cudaStream_t st) {
    int cap = block_capacity(dt), bb = block_bytes(dt);
    IFB_CHECK_ARG(is_quantized(dt) && cap > 1, "Dequantize: element type %d is not supported", dt);
    IFB_CHECK_ARG(cx % cap == 0, "The number of columns should be a multiple of %d", cap);
    if (rows <= 0) return IFB_OK;
    int64_t nblocks = rows * (cx / cap);
    dequantize_blocks_kernel<<<(unsigned)((nblocks + 127) / 128), 128, 0, st>>>(
        dt, (__half *)dst, (const uint8_t *)src, nblocks, cap, bb);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

}  // namespace ifb
