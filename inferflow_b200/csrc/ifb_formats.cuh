// Inferflow block-quantization formats: the on-device storage contract.
//
// Byte layouts are the reference's (src/common/quant_types.h:11-174); element
// type ids are the reference's ElementType values (src/tensor/tensor_common.h:15-42)
// so a binding passes `(int)tensor.data_type` straight through.
//
// Device decoders below produce, for every format, the block's integer codes
// packed four-per-32-bit-word in the order the dp4a kernels consume them
// (the job of Quantization::GetInt4, src/common/quantization.h:160,227,401,506,714,809),
// plus the block's fp16 `scale` and `base`:  value = code * scale + base.
#pragma once

#include "ifb_common.cuh"

namespace ifb {

enum DType : int {
    DT_F32 = 0,
    DT_F16 = 1,
    DT_BF16 = 2,
    DT_I32 = 3,
    DT_I16 = 4,
    DT_Q8_GL = 5,
    DT_Q8_LOG = 6,
    DT_Q8_B32T1 = 7,
    DT_Q8_B32T2 = 8,
    DT_Q6_B64T1 = 9,
    DT_Q5_B64T1 = 10,
    DT_Q5_B32T1 = 11,
    DT_Q4_B16 = 12,
    DT_Q4_B32T1A = 13,
    DT_Q4_B32T1B = 14,
    DT_Q4_B32T2 = 15,
    DT_Q4_B32P8 = 16,
    DT_Q4_B64T1 = 17,
    DT_Q3H_B64T1 = 18,
    DT_Q3_B32T1A = 19,
    DT_Q3_B32T1B = 20,
    DT_Q2_B32T1A = 21,
    DT_Q2_B32T1B = 22,
    DT_INVALID = 23,
};

// block capacity (elements) and size (bytes); 0 when the type has no block kernels here
__host__ __device__ inline int block_capacity(int dt) {
    switch (dt) {
        case DT_F16: return 1;
        case DT_F32: return 1;
        case DT_Q8_B32T1: case DT_Q8_B32T2: case DT_Q5_B32T1: case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: case DT_Q3_B32T1A: case DT_Q3_B32T1B: case DT_Q2_B32T1A:
        case DT_Q2_B32T1B: return 32;
        case DT_Q6_B64T1: case DT_Q5_B64T1: case DT_Q4_B64T1: case DT_Q3H_B64T1: return 64;
        case DT_Q4_B16: return 16;
        default: return 0;
    }
}
__host__ __device__ inline int block_bytes(int dt) {
    switch (dt) {
        case DT_F16: return 2;
        case DT_F32: return 4;
        case DT_Q8_B32T1: return 36;
        case DT_Q8_B32T2: return 34;
        case DT_Q6_B64T1: return 52;
        case DT_Q5_B64T1: return 44;
        case DT_Q5_B32T1: return 24;
        case DT_Q4_B16: return 10;
        case DT_Q4_B32T1A: case DT_Q4_B32T1B: return 20;
        case DT_Q4_B64T1: return 36;
        case DT_Q3H_B64T1: return 32;
        case DT_Q3_B32T1A: case DT_Q3_B32T1B: return 16;
        case DT_Q2_B32T1A: case DT_Q2_B32T1B: return 12;
        default: return 0;
    }
}
// bytes of one row of cx elements (reference: TensorCommon::ByteCount, tensor_common.cc:162-167)
__host__ __device__ inline int64_t row_bytes(int dt, int64_t cx) {
    int cap = block_capacity(dt);
    if (cap == 0) return -1;
    return (cx / cap) * (int64_t)block_bytes(dt);
}
__host__ __device__ inline bool is_quantized(int dt) { return dt > DT_BF16; }

// ---------------------------------------------------------------------------
// Format traits for the dp4a ("AX8") GEMV: weights that meet a Q8_B32T2 activation
// (reference default path: GetUseFullQuantGemv, inference_worker.cc:2707-2730).
//
//  CAP       elements per weight block (32 or 64)
//  BYTES     block size
//  XSPLIT    activation codes are staged in "nibble-split" order per group of 8:
//            (x0,x2,x4,x6)(x1,x3,x5,x7), so that a word of 8 packed nibbles needs only
//            AND / SHR+AND before two dp4a (no byte permute).
//  HAS_BASE  value = q*scale + base (false: Q8_B32T2, value = q*scale, signed codes)
//  decode(): blk points at the block in shared memory (4-byte aligned unless noted),
//            fills codes[CAP/4] and returns scale/base.
// ---------------------------------------------------------------------------
struct FmtQ4_B32T1 {
    static constexpr int CAP = 32, BYTES = 20, NW = 8;
    static constexpr bool XSPLIT = true, HAS_BASE = true, SIGNED = false;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(blk);
        uint32_t h = w[0];
        base = h2f((uint16_t)(h & 0xffffu));
        scale = h2f((uint16_t)(h >> 16));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t d = w[1 + i];
            codes[2 * i] = d & 0x0f0f0f0fu;
            codes[2 * i + 1] = (d >> 4) & 0x0f0f0f0fu;
        }
    }
};

struct FmtQ4_B64T1 {
    static constexpr int CAP = 64, BYTES = 36, NW = 16;
    static constexpr bool XSPLIT = true, HAS_BASE = true, SIGNED = false;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(blk);
        uint32_t h = w[0];
        base = h2f((uint16_t)(h & 0xffffu));
        scale = h2f((uint16_t)(h >> 16));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t d = w[1 + i];
            codes[2 * i] = d & 0x0f0f0f0fu;
            codes[2 * i + 1] = (d >> 4) & 0x0f0f0f0fu;
        }
    }
};

// Q8_B32T2: fp16 scale + 32 int8; 34-byte blocks are only 2-byte aligned.
struct FmtQ8_B32T2 {
    static constexpr int CAP = 32, BYTES = 34, NW = 8;
    static constexpr bool XSPLIT = false, HAS_BASE = false, SIGNED = true;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        uint32_t addr = smem_u32(blk);
        scale = h2f(*reinterpret_cast<const uint16_t *>(blk));
        base = 0.f;
        // data bytes start at blk+2; realign with funnel shifts
        uint32_t a = (addr + 2u) & ~3u;
        uint32_t sh = ((addr + 2u) & 3u) * 8u;  // 0 or 16
        const uint32_t *w =
            reinterpret_cast<const uint32_t *>(blk + 2 - ((addr + 2u) & 3u));
        (void)a;
        uint32_t prev = w[0];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t next = w[i + 1];  // last one may read 4 bytes past the block: slots are padded
            codes[i] = __funnelshift_r(prev, next, sh);
            prev = next;
        }
    }
};

// Q6_B64T1: base, scale, data_h[16] (2 high bits x4 per byte), data[32] (low nibbles)
struct FmtQ6_B64T1 {
    static constexpr int CAP = 64, BYTES = 52, NW = 16;
    static constexpr bool XSPLIT = false, HAS_BASE = true, SIGNED = false;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(blk);
        uint32_t h = w[0];
        base = h2f((uint16_t)(h & 0xffffu));
        scale = h2f((uint16_t)(h >> 16));
        // words 1..4: data_h (16 bytes), words 5..12: data (32 bytes)
#pragma unroll
        for (int j = 0; j < 8; j++) {      // data word j -> groups 2j, 2j+1 (8 elements)
            uint32_t d = w[5 + j];
            uint32_t hh = (w[1 + j / 2] >> ((j & 1) * 16)) & 0xffffu;  // 2 high-bytes for the 2 groups
            // group g = 2j: nibbles 0..3 of d (low 16 bits), high bits byte hh&0xff
            uint32_t lo16 = d & 0xffffu, hi16 = d >> 16;
            uint32_t h0 = hh & 0xffu, h1 = hh >> 8;
            // spread 4 nibbles of a 16-bit value into 4 bytes
            uint32_t n0 = (lo16 & 0x000fu) | ((lo16 & 0x00f0u) << 4) | ((lo16 & 0x0f00u) << 8) |
                          ((lo16 & 0xf000u) << 12);
            uint32_t n1 = (hi16 & 0x000fu) | ((hi16 & 0x00f0u) << 4) | ((hi16 & 0x0f00u) << 8) |
                          ((hi16 & 0xf000u) << 12);
            uint32_t b0 = ((h0 & 0x03u) << 4) | ((h0 & 0x0cu) << 10) | ((h0 & 0x30u) << 16) |
                          ((h0 & 0xc0u) << 22);
            uint32_t b1 = ((h1 & 0x03u) << 4) | ((h1 & 0x0cu) << 10) | ((h1 & 0x30u) << 16) |
                          ((h1 & 0xc0u) << 22);
            codes[2 * j] = n0 | b0;
            codes[2 * j + 1] = n1 | b1;
        }
    }
};

// Q5_B64T1: base, scale, data_h[8] (bit j of data_h[r] = bit 4 of q[8r+j]), data[32]
struct FmtQ5_B64T1 {
    static constexpr int CAP = 64, BYTES = 44, NW = 16;
    static constexpr bool XSPLIT = false, HAS_BASE = true, SIGNED = false;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(blk);
        uint32_t h = w[0];
        base = h2f((uint16_t)(h & 0xffffu));
        scale = h2f((uint16_t)(h >> 16));
        // words 1..2: data_h (8 bytes), words 3..10: data
#pragma unroll
        for (int j = 0; j < 8; j++) {  // data word j: elements 8j..8j+7, high bits = data_h[j]
            uint32_t d = w[3 + j];
            uint32_t hb = (w[1 + j / 4] >> ((j & 3) * 8)) & 0xffu;
            uint32_t lo16 = d & 0xffffu, hi16 = d >> 16;
            uint32_t n0 = (lo16 & 0x000fu) | ((lo16 & 0x00f0u) << 4) | ((lo16 & 0x0f00u) << 8) |
                          ((lo16 & 0xf000u) << 12);
            uint32_t n1 = (hi16 & 0x000fu) | ((hi16 & 0x00f0u) << 4) | ((hi16 & 0x0f00u) << 8) |
                          ((hi16 & 0xf000u) << 12);
            uint32_t b0 = ((hb & 0x01u) << 4) | ((hb & 0x02u) << 11) | ((hb & 0x04u) << 18) |
                          ((hb & 0x08u) << 25);
            uint32_t h4 = hb >> 4;
            uint32_t b1 = ((h4 & 0x01u) << 4) | ((h4 & 0x02u) << 11) | ((h4 & 0x04u) << 18) |
                          ((h4 & 0x08u) << 25);
            codes[2 * j] = n0 | b0;
            codes[2 * j + 1] = n1 | b1;
        }
    }
};

// Q3H_B64T1 ("3.5 bit"): 11 levels, two codes per 7-bit pair P = q0 + 11*q1.
// base, scale, data_h[4], data_m[8], data[16]
struct FmtQ3H_B64T1 {
    static constexpr int CAP = 64, BYTES = 32, NW = 16;
    static constexpr bool XSPLIT = false, HAS_BASE = true, SIGNED = false;
    __device__ static __forceinline__ void decode(const uint8_t *blk, uint32_t *codes,
                                                  float &scale, float &base) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(blk);
        uint32_t h = w[0];
        base = h2f((uint16_t)(h & 0xffffu));
        scale = h2f((uint16_t)(h >> 16));
        uint32_t dh = w[1];                 // data_h[0..3]
        // data_m: words 2..3 ; data: words 4..7
#pragma unroll
        for (int r = 0; r < 8; r++) {       // group r: 4 pairs = 8 elements
            uint32_t u16 = (w[4 + r / 2] >> ((r & 1) * 16)) & 0xffffu;
            uint32_t m8 = (w[2 + r / 4] >> ((r & 3) * 8)) & 0xffu;
            uint32_t h8 = (dh >> (r * 4)) & 0x0fu;  // data_h[r/2] low (r even) / high (r odd) nibble
            uint32_t p0 = (u16 & 0x000fu) | ((m8 & 0x03u) << 4) | ((h8 & 0x01u) << 6);
            uint32_t p1 = ((u16 & 0x00f0u) >> 4) | ((m8 & 0x0cu) << 2) | ((h8 & 0x02u) << 5);
            uint32_t p2 = ((u16 & 0x0f00u) >> 8) | (m8 & 0x30u) | ((h8 & 0x04u) << 4);
            uint32_t p3 = ((u16 & 0xf000u) >> 12) | ((m8 & 0xc0u) >> 2) | ((h8 & 0x08u) << 3);
            // P/11 for P in 0..127 : (P*187)>>11 is exact on that range (checked exhaustively in tests)
            uint32_t d0 = (p0 * 187u) >> 11, d1 = (p1 * 187u) >> 11, d2 = (p2 * 187u) >> 11,
                     d3 = (p3 * 187u) >> 11;
            uint32_t m0 = p0 - 11u * d0, m1 = p1 - 11u * d1, m2 = p2 - 11u * d2,
                     m3 = p3 - 11u * d3;
            codes[2 * r] = m0 | (d0 << 8) | (m1 << 16) | (d1 << 24);
            codes[2 * r + 1] = m2 | (d2 << 8) | (m3 << 16) | (d3 << 24);
        }
    }
};

// ---------------------------------------------------------------------------
// Generic per-element dequantization of one block to float (fp16-x GEMV path,
// TensorOpr::Dequantize, GEMM operand expansion). Mirrors Quantization::Dequantize*
// (quantization.h) -- value = q*scale + base in fp32; callers round to half where
// the reference stores half.
// blk must be readable byte-wise (global or shared).
// ---------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void dequant_block(int dt, const uint8_t *blk, F &&emit) {
    // emit(index_in_block, float_value)
    switch (dt) {
        case DT_Q8_B32T1: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            for (int i = 0; i < 32; i++) emit(i, (float)blk[4 + i] * scale + base);
        } break;
        case DT_Q8_B32T2: {
            float scale = h2f(*(const uint16_t *)(blk));
            for (int i = 0; i < 32; i++) emit(i, (float)(int)(int8_t)blk[2 + i] * scale);
        } break;
        case DT_Q6_B64T1: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            const uint8_t *dh = blk + 4, *d = blk + 20;
            for (int g = 0; g < 16; g++) {
                uint32_t qh = dh[g];
                uint32_t qd = d[2 * g] | ((uint32_t)d[2 * g + 1] << 8);
                for (int k = 0; k < 4; k++) {
                    int q = (int)((qd >> (4 * k)) & 0xf) | (int)(((qh >> (2 * k)) & 3) << 4);
                    emit(4 * g + k, (float)q * scale + base);
                }
            }
        } break;
        case DT_Q5_B64T1: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            const uint8_t *dh = blk + 4, *d = blk + 12;
            for (int i = 0; i < 64; i++) {
                int lo = (d[i / 2] >> ((i & 1) * 4)) & 0xf;
                int hi = (dh[i / 8] >> (i & 7)) & 1;
                emit(i, (float)(lo | (hi << 4)) * scale + base);
            }
        } break;
        case DT_Q5_B32T1: {
            float scale = h2f(*(const uint16_t *)(blk)), base = h2f(*(const uint16_t *)(blk + 2));
            uint32_t qh = blk[4] | ((uint32_t)blk[5] << 8) | ((uint32_t)blk[6] << 16) |
                          ((uint32_t)blk[7] << 24);
            const uint8_t *d = blk + 8;
            for (int i = 0; i < 16; i++) {
                int x0 = (d[i] & 0xf) | (int)(((qh >> i) & 1) << 4);
                int x1 = (d[i] >> 4) | (int)(((qh >> (i + 16)) & 1) << 4);
                emit(i, (float)x0 * scale + base);
                emit(i + 16, (float)x1 * scale + base);
            }
        } break;
        case DT_Q4_B16: {
            float scale = (float)blk[1] / 1000;
            float base = (int)(uint32_t)blk[0] / 100.0f - 1.0f;
            for (int i = 0; i < 8; i++) {
                emit(2 * i, (float)(blk[2 + i] & 0xf) * scale + base);
                emit(2 * i + 1, (float)(blk[2 + i] >> 4) * scale + base);
            }
        } break;
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            for (int i = 0; i < 16; i++) {
                emit(2 * i, (float)(blk[4 + i] & 0xf) * scale + base);
                emit(2 * i + 1, (float)(blk[4 + i] >> 4) * scale + base);
            }
        } break;
        case DT_Q4_B64T1: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            for (int i = 0; i < 32; i++) {
                emit(2 * i, (float)(blk[4 + i] & 0xf) * scale + base);
                emit(2 * i + 1, (float)(blk[4 + i] >> 4) * scale + base);
            }
        } break;
        case DT_Q3H_B64T1: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            const uint8_t *dh = blk + 4, *dm = blk + 8, *d = blk + 16;
            for (int p = 0; p < 32; p++) {  // pair p -> elements 2p, 2p+1
                int r = p / 4, k = p & 3;
                int lo = (d[2 * r + k / 2] >> ((k & 1) * 4)) & 0xf;
                int mid = (dm[r] >> (2 * k)) & 3;
                int hi = (dh[r / 2] >> ((r & 1) * 4 + k)) & 1;
                int P = lo | (mid << 4) | (hi << 6);
                emit(2 * p, (float)(P % 11) * scale + base);
                emit(2 * p + 1, (float)(P / 11) * scale + base);
            }
        } break;
        case DT_Q3_B32T1A:
        case DT_Q3_B32T1B: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            const uint8_t *hd = blk + 4, *d = blk + 8;
            for (int i = 0; i < 32; i++) {
                int lo = (d[i / 4] >> ((i & 3) * 2)) & 3;
                int hi = (hd[i / 8] >> (i & 7)) & 1;
                emit(i, (float)(lo | (hi << 2)) * scale + base);
            }
        } break;
        case DT_Q2_B32T1A:
        case DT_Q2_B32T1B: {
            float base = h2f(*(const uint16_t *)(blk)), scale = h2f(*(const uint16_t *)(blk + 2));
            for (int i = 0; i < 32; i++)
                emit(i, (float)((blk[4 + i / 4] >> ((i & 3) * 2)) & 3) * scale + base);
        } break;
        default: break;
    }
}

}  // namespace ifb
