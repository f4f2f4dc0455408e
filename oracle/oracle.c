/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. See oracle.h.
 *
 * Plain-C restatement of the reference algorithms on the decode hot path. Nothing
 * here is copied from the reference: the formulas are restated from the file:line
 * cited at each function and pinned against the reference compiled in this
 * container (tests/test_oracle_vs_ref.py, tests/golden/).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ fp16 */
uint16_t orc_f32_to_f16(float v) {
    _Float16 h = (_Float16)v; /* IEEE RN-even */
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
float orc_f16_to_f32(uint16_t b) {
    _Float16 h;
    memcpy(&h, &b, 2);
    return (float)h;
}
void orc_round_f16(float *v, int64_t n) {
    for (int64_t i = 0; i < n; i++) v[i] = (float)(_Float16)v[i];
}
static inline void put16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v & 0xff); p[1] = (uint8_t)(v >> 8); }
static inline uint16_t get16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ------------------------------------------------------------------ formats
 * quant_types.h:11-174 (sizes verified against sizeof in tests) */
int orc_block_capacity(int dt) {
    switch (dt) {
        case ORC_F32: case ORC_F16: return 1;
        case ORC_Q8_B32T1: case ORC_Q8_B32T2: case ORC_Q5_B32T1: case ORC_Q4_B32T1A:
        case ORC_Q4_B32T1B: case ORC_Q3_B32T1A: case ORC_Q3_B32T1B: case ORC_Q2_B32T1A:
        case ORC_Q2_B32T1B: return 32;
        case ORC_Q6_B64T1: case ORC_Q5_B64T1: case ORC_Q4_B64T1: case ORC_Q3H_B64T1: return 64;
        case ORC_Q4_B16: return 16;
        default: return 0;
    }
}
int orc_block_bytes(int dt) {
    switch (dt) {
        case ORC_F32: return 4;
        case ORC_F16: return 2;
        case ORC_Q8_B32T1: return 36;
        case ORC_Q8_B32T2: return 34;
        case ORC_Q6_B64T1: return 52;
        case ORC_Q5_B64T1: return 44;
        case ORC_Q5_B32T1: return 24;
        case ORC_Q4_B16: return 10;
        case ORC_Q4_B32T1A: case ORC_Q4_B32T1B: return 20;
        case ORC_Q4_B64T1: return 36;
        case ORC_Q3H_B64T1: return 32;
        case ORC_Q3_B32T1A: case ORC_Q3_B32T1B: return 16;
        case ORC_Q2_B32T1A: case ORC_Q2_B32T1B: return 12;
        default: return 0;
    }
}
int64_t orc_row_bytes(int dt, int64_t cx) {
    int cap = orc_block_capacity(dt);
    return cap ? (cx / cap) * (int64_t)orc_block_bytes(dt) : -1;
}

/* value range in element order, fp32 (quantization.h:71-87) */
static void value_range(const float *a, int n, float *mn, float *mx) {
    *mn = a[0]; *mx = a[0];
    for (int i = 1; i < n; i++) {
        if (*mn > a[i]) *mn = a[i];
        if (*mx < a[i]) *mx = a[i];
    }
}

/* Linear min/max quantizer shared by most formats:
 *   scale = (max-min)/div ; inv = scale >= 1e-5 ? 1/scale : 0 ;
 *   mode 0: base=min,          q = (uint32)(v*inv + 0.5)         (round)
 *   mode 1: base=min+scale/2,  q = (uint32)(v*inv + 0.0001)      (trunc, "B" variants)
 * FMA: the reference as built contracts `(v-min)*inv + 0.5f` into one fused
 * multiply-add on BOTH flavours (g++ -O2 -mfma defaults to -ffp-contract=fast for C++,
 * checked in the disassembly of oracle/_ref; nvcc defaults to --fmad=true), so the
 * restatement uses fmaf() explicitly; same for q*scale+base in the dequantizers. */
typedef struct { float mn, scale, inv, base; } qparams_t;
static qparams_t lin_params(const float *src, int n, float div, int mode) {
    qparams_t p; float mx;
    value_range(src, n, &p.mn, &mx);
    p.scale = (mx - p.mn) / div;
    p.inv = p.scale >= 0.00001f ? (1.0f / p.scale) : 0.0f;
    p.base = mode ? fmaf(0.5f, p.scale, p.mn) : p.mn;
    return p;
}
static inline uint32_t lin_q(float v, const qparams_t *p, int mode, uint32_t maxq) {
    float u = fmaf(v - p->mn, p->inv, mode ? 0.0001f : 0.5f);
    uint32_t q = (uint32_t)u;
    return q > maxq ? maxq : q;
}

/* ---- per-format block quantizers -------------------------------------- */
static void q_q8_b32t1(uint8_t *b, const float *s) { /* quantization.h:110-154 */
    qparams_t p = lin_params(s, 32, 255.0f, 0);
    put16(b, orc_f32_to_f16(p.mn)); put16(b + 2, orc_f32_to_f16(p.scale));
    for (int i = 0; i < 32; i++) b[4 + i] = (uint8_t)lin_q(s[i], &p, 0, 255);
}
static void q_q8_b32t2_dev(uint8_t *b, const float *s) { /* tensor_quant.h:45-82 (alg 2) */
    float m = 0;
    for (int i = 0; i < 32; i++) { float a = fabsf(s[i]); if (a > m) m = a; }
    float scale = m / 127.0f;
    put16(b, orc_f32_to_f16(scale));
    for (int i = 0; i < 32; i++) {
        int q = scale <= 0.000001f ? 0 : (int)roundf(s[i] / scale);
        if (q < -128) q = -128; if (q > 127) q = 127;
        b[2 + i] = (uint8_t)(int8_t)q;
    }
}
static void q_q8_b32t2_host(uint8_t *b, const float *s) { /* quantization.h:186-221 */
    float mn, mx; value_range(s, 32, &mn, &mx);
    float m1 = fabsf(mx), m2 = fabsf(mn), m0 = m1 > m2 ? m1 : m2;
    float scale = m0 / 127.0f;
    float inv = scale >= 0.00001f ? (1.0f / scale) : 0.0f;
    put16(b, orc_f32_to_f16(scale));
    for (int i = 0; i < 32; i++) {
        volatile float qf = s[i] * inv;
        int q = (int)round((double)qf);
        if (q > 127) q = 127; if (q < -128) q = -128;
        b[2 + i] = (uint8_t)(int8_t)q;
    }
}
static void q_q6_b64t1(uint8_t *b, const float *s) { /* quantization.h:268-318: 62 steps, clamp 63 */
    qparams_t p = lin_params(s, 64, 62.0f, 0);
    put16(b, orc_f32_to_f16(p.mn)); put16(b + 2, orc_f32_to_f16(p.scale));
    uint8_t *dh = b + 4, *d = b + 20;
    for (int r = 0; r < 16; r++) {
        uint32_t q[4];
        for (int k = 0; k < 4; k++) q[k] = lin_q(s[4 * r + k], &p, 0, 63);
        dh[r] = (uint8_t)((q[0] >> 4) | ((q[1] >> 4) << 2) | ((q[2] >> 4) << 4) | ((q[3] >> 4) << 6));
        d[2 * r] = (uint8_t)((q[0] & 15) | ((q[1] & 15) << 4));
        d[2 * r + 1] = (uint8_t)((q[2] & 15) | ((q[3] & 15) << 4));
    }
}
static void q_q5_b64t1(uint8_t *b, const float *s) { /* quantization.h:446-500: 30 steps, clamp 31 */
    qparams_t p = lin_params(s, 64, 30.0f, 0);
    put16(b, orc_f32_to_f16(p.mn)); put16(b + 2, orc_f32_to_f16(p.scale));
    uint8_t *dh = b + 4, *d = b + 12;
    for (int r = 0; r < 8; r++) {
        uint32_t q[8], h = 0;
        for (int k = 0; k < 8; k++) { q[k] = lin_q(s[8 * r + k], &p, 0, 31); h |= ((q[k] >> 4) & 1) << k; }
        dh[r] = (uint8_t)h;
        for (int k = 0; k < 4; k++) d[4 * r + k] = (uint8_t)((q[2 * k] & 15) | ((q[2 * k + 1] & 15) << 4));
    }
}
static void q_q5_b32t1(uint8_t *b, const float *s) { /* quantization.h:348-395: /31, round, NO clamp */
    qparams_t p = lin_params(s, 32, 31.0f, 0);
    put16(b, orc_f32_to_f16(p.scale)); put16(b + 2, orc_f32_to_f16(p.mn)); /* scale first */
    uint32_t qh = 0;
    for (int r = 0; r < 16; r++) {
        uint32_t q1 = lin_q(s[r], &p, 0, 0xffffffffu), q2 = lin_q(s[r + 16], &p, 0, 0xffffffffu);
        b[8 + r] = (uint8_t)((q1 & 15) | ((q2 & 15) << 4));
        qh |= ((q1 & 16) >> 4) << r;
        qh |= ((q2 & 16) >> 4) << (r + 16);
    }
    b[4] = (uint8_t)qh; b[5] = (uint8_t)(qh >> 8); b[6] = (uint8_t)(qh >> 16); b[7] = (uint8_t)(qh >> 24);
}
static void q_q4_b16(uint8_t *b, const float *s) { /* quantization.h:43-68,657-708 */
    float mn, mx; value_range(s, 16, &mn, &mx);
    /* AdjustBase: u8 = uint8(base*100 + 100.01) [double arithmetic], DecodeBase */
    uint8_t u8 = (uint8_t)((double)(mn * 100) + 100.01);
    mn = (int)(uint32_t)u8 / 100.0f - 1.0f;
    float scale = (mx - mn) / 15.0f;
    float inv = scale >= 0.00001f ? (1.0f / scale) : 0.0f;
    b[0] = (uint8_t)fmaf(mn, 100.0f, 100.5f);
    b[1] = (uint8_t)fmaf(scale, 1000.0f, 0.5f);
    qparams_t p = { mn, scale, inv, mn };
    for (int r = 0; r < 8; r++) {
        uint32_t q1 = lin_q(s[2 * r], &p, 0, 15), q2 = lin_q(s[2 * r + 1], &p, 0, 15);
        b[2 + r] = (uint8_t)(q1 | (q2 << 4));
    }
}
static void q_q4_b32t1(uint8_t *b, const float *s, int mode) { /* A: :535-587 (/15), B: :589-631 (/16) */
    qparams_t p = lin_params(s, 32, mode ? 16.0f : 15.0f, mode);
    put16(b, orc_f32_to_f16(p.base)); put16(b + 2, orc_f32_to_f16(p.scale));
    for (int r = 0; r < 16; r++) {
        uint32_t q1 = lin_q(s[2 * r], &p, mode, 15), q2 = lin_q(s[2 * r + 1], &p, mode, 15);
        b[4 + r] = (uint8_t)(q1 | (q2 << 4));
    }
}
static void q_q4_b64t1(uint8_t *b, const float *s) { /* quantization.h:757-803: 14 steps, clamp 15 */
    qparams_t p = lin_params(s, 64, 14.0f, 0);
    put16(b, orc_f32_to_f16(p.mn)); put16(b + 2, orc_f32_to_f16(p.scale));
    for (int r = 0; r < 32; r++) {
        uint32_t q1 = lin_q(s[2 * r], &p, 0, 15), q2 = lin_q(s[2 * r + 1], &p, 0, 15);
        b[4 + r] = (uint8_t)(q1 | (q2 << 4));
    }
}
static void q_q3h_b64t1(uint8_t *b, const float *s) { /* quantization.h:854-926: 11 levels, pairs base 11 */
    qparams_t p = lin_params(s, 64, 10.0f, 0);
    put16(b, orc_f32_to_f16(p.mn)); put16(b + 2, orc_f32_to_f16(p.scale));
    uint8_t *dh = b + 4, *dm = b + 8, *d = b + 16;
    memset(dh, 0, 4);
    for (int r = 0; r < 8; r++) {
        int P[4];
        for (int k = 0; k < 4; k++) {
            /* (int)(v+0.5f), clamp [0,10] */
            float u0 = fmaf(s[8 * r + 2 * k] - p.mn, p.inv, 0.5f);
            float u1 = fmaf(s[8 * r + 2 * k + 1] - p.mn, p.inv, 0.5f);
            int q0 = (int)u0, q1 = (int)u1;
            if (q0 < 0) q0 = 0; if (q0 > 10) q0 = 10;
            if (q1 < 0) q1 = 0; if (q1 > 10) q1 = 10;
            P[k] = q0 + 11 * q1;
        }
        d[2 * r] = (uint8_t)((P[0] & 15) | ((P[1] & 15) << 4));
        d[2 * r + 1] = (uint8_t)((P[2] & 15) | ((P[3] & 15) << 4));
        dm[r] = (uint8_t)(((P[0] >> 4) & 3) | (((P[1] >> 4) & 3) << 2) | (((P[2] >> 4) & 3) << 4) |
                          (((P[3] >> 4) & 3) << 6));
        int hn = ((P[0] >> 6) & 1) | (((P[1] >> 6) & 1) << 1) | (((P[2] >> 6) & 1) << 2) |
                 (((P[3] >> 6) & 1) << 3);
        dh[r / 2] |= (uint8_t)(hn << ((r & 1) * 4));
    }
}
static void q_q3_b32t1(uint8_t *b, const float *s, int mode) { /* A: :964-1014 (/7), B: :1016-1068 (/8) */
    qparams_t p = lin_params(s, 32, mode ? 8.0f : 7.0f, mode);
    put16(b, orc_f32_to_f16(p.base)); put16(b + 2, orc_f32_to_f16(p.scale));
    uint8_t *hd = b + 4, *d = b + 8;
    for (int r = 0; r < 4; r++) {
        uint32_t q[8], h = 0, lo = 0, hi = 0;
        for (int k = 0; k < 8; k++) { q[k] = lin_q(s[8 * r + k], &p, mode, 7); h |= ((q[k] >> 2) & 1) << k; }
        for (int k = 0; k < 4; k++) { lo |= (q[k] & 3) << (2 * k); hi |= (q[4 + k] & 3) << (2 * k); }
        d[2 * r] = (uint8_t)lo; d[2 * r + 1] = (uint8_t)hi; hd[r] = (uint8_t)h;
    }
}
static void q_q2_b32t1(uint8_t *b, const float *s, int mode) { /* A: :1098-1146 (/3), B: :1148-1194 (/4) */
    qparams_t p = lin_params(s, 32, mode ? 4.0f : 3.0f, mode);
    put16(b, orc_f32_to_f16(p.base)); put16(b + 2, orc_f32_to_f16(p.scale));
    for (int r = 0; r < 8; r++) {
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) v |= lin_q(s[4 * r + k], &p, mode, 3) << (2 * k);
        b[4 + r] = (uint8_t)v;
    }
}

int orc_quantize_row(int dt, uint8_t *dst, const float *src, int64_t n) {
    int cap = orc_block_capacity(dt), bb = orc_block_bytes(dt);
    if (cap <= 1) return -1;
    for (int64_t i = 0; i < n / cap; i++) {
        uint8_t *b = dst + i * bb; const float *s = src + i * cap;
        switch (dt) {
            case ORC_Q8_B32T1: q_q8_b32t1(b, s); break;
            case ORC_Q8_B32T2: q_q8_b32t2_dev(b, s); break;
            case ORC_Q6_B64T1: q_q6_b64t1(b, s); break;
            case ORC_Q5_B64T1: q_q5_b64t1(b, s); break;
            case ORC_Q5_B32T1: q_q5_b32t1(b, s); break;
            case ORC_Q4_B16: q_q4_b16(b, s); break;
            case ORC_Q4_B32T1A: q_q4_b32t1(b, s, 0); break;
            case ORC_Q4_B32T1B: q_q4_b32t1(b, s, 1); break;
            case ORC_Q4_B64T1: q_q4_b64t1(b, s); break;
            case ORC_Q3H_B64T1: q_q3h_b64t1(b, s); break;
            case ORC_Q3_B32T1A: q_q3_b32t1(b, s, 0); break;
            case ORC_Q3_B32T1B: q_q3_b32t1(b, s, 1); break;
            case ORC_Q2_B32T1A: q_q2_b32t1(b, s, 0); break;
            case ORC_Q2_B32T1B: q_q2_b32t1(b, s, 1); break;
            default: return -1;
        }
    }
    return 0;
}
int orc_quantize_row_q8t2_host(uint8_t *dst, const float *src, int64_t n) {
    for (int64_t i = 0; i < n / 32; i++) q_q8_b32t2_host(dst + i * 34, src + i * 32);
    return 0;
}

/* ---- decode: codes of one block in element order + scale/base ----------- */
int orc_block_codes(int dt, int32_t *c, float *scale, float *base, const uint8_t *b) {
    *scale = 0; *base = 0;
    switch (dt) {
        case ORC_Q8_B32T1:
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            for (int i = 0; i < 32; i++) c[i] = b[4 + i];
            return 32;
        case ORC_Q8_B32T2:
            *scale = orc_f16_to_f32(get16(b));
            for (int i = 0; i < 32; i++) c[i] = (int8_t)b[2 + i];
            return 32;
        case ORC_Q6_B64T1: {
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            const uint8_t *dh = b + 4, *d = b + 20;
            for (int i = 0; i < 64; i++) {
                int lo = (d[i / 2] >> ((i & 1) * 4)) & 15;
                int hi = (dh[i / 4] >> ((i & 3) * 2)) & 3;
                c[i] = lo | (hi << 4);
            }
            return 64;
        }
        case ORC_Q5_B64T1: {
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            const uint8_t *dh = b + 4, *d = b + 12;
            for (int i = 0; i < 64; i++)
                c[i] = ((d[i / 2] >> ((i & 1) * 4)) & 15) | (((dh[i / 8] >> (i & 7)) & 1) << 4);
            return 64;
        }
        case ORC_Q5_B32T1: {
            *scale = orc_f16_to_f32(get16(b)); *base = orc_f16_to_f32(get16(b + 2));
            uint32_t qh = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
            for (int i = 0; i < 16; i++) {
                c[i] = (b[8 + i] & 15) | (((qh >> i) & 1) << 4);
                c[i + 16] = (b[8 + i] >> 4) | (((qh >> (i + 16)) & 1) << 4);
            }
            return 32;
        }
        case ORC_Q4_B16:
            *scale = (float)b[1] / 1000; *base = (int)(uint32_t)b[0] / 100.0f - 1.0f;
            for (int i = 0; i < 8; i++) { c[2 * i] = b[2 + i] & 15; c[2 * i + 1] = b[2 + i] >> 4; }
            return 16;
        case ORC_Q4_B32T1A: case ORC_Q4_B32T1B:
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            for (int i = 0; i < 16; i++) { c[2 * i] = b[4 + i] & 15; c[2 * i + 1] = b[4 + i] >> 4; }
            return 32;
        case ORC_Q4_B64T1:
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            for (int i = 0; i < 32; i++) { c[2 * i] = b[4 + i] & 15; c[2 * i + 1] = b[4 + i] >> 4; }
            return 64;
        case ORC_Q3H_B64T1: {
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            const uint8_t *dh = b + 4, *dm = b + 8, *d = b + 16;
            for (int p = 0; p < 32; p++) {
                int r = p / 4, k = p & 3;
                int P = ((d[2 * r + k / 2] >> ((k & 1) * 4)) & 15) | (((dm[r] >> (2 * k)) & 3) << 4) |
                        (((dh[r / 2] >> ((r & 1) * 4 + k)) & 1) << 6);
                c[2 * p] = P % 11; c[2 * p + 1] = P / 11;
            }
            return 64;
        }
        case ORC_Q3_B32T1A: case ORC_Q3_B32T1B: {
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            const uint8_t *hd = b + 4, *d = b + 8;
            for (int i = 0; i < 32; i++)
                c[i] = ((d[i / 4] >> ((i & 3) * 2)) & 3) | (((hd[i / 8] >> (i & 7)) & 1) << 2);
            return 32;
        }
        case ORC_Q2_B32T1A: case ORC_Q2_B32T1B:
            *base = orc_f16_to_f32(get16(b)); *scale = orc_f16_to_f32(get16(b + 2));
            for (int i = 0; i < 32; i++) c[i] = (b[4 + i / 4] >> ((i & 3) * 2)) & 3;
            return 32;
        default: return -1;
    }
}

int orc_dequantize_row(int dt, float *dst, const uint8_t *src, int64_t n) {
    int cap = orc_block_capacity(dt), bb = orc_block_bytes(dt);
    if (cap <= 1) return -1;
    int32_t c[64]; float sc, ba;
    for (int64_t i = 0; i < n / cap; i++) {
        if (orc_block_codes(dt, c, &sc, &ba, src + i * bb) < 0) return -1;
        for (int k = 0; k < cap; k++) {
            dst[i * cap + k] = fmaf((float)c[k], sc, ba); /* q*scale + base, contracted */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ GEMV */
int orc_gemv_ax8(int dt, double *y, const uint8_t *w, const uint8_t *xq, int64_t cy, int64_t cx) {
    int cap = orc_block_capacity(dt), bb = orc_block_bytes(dt);
    if (cap != 32 && cap != 64) return -1;
    int64_t nb = cx / cap, nxb = cx / 32;
    float *sx = (float *)malloc(sizeof(float) * nxb);
    int32_t *qx = (int32_t *)malloc(sizeof(int32_t) * cx);
    for (int64_t b = 0; b < nxb; b++) {
        sx[b] = orc_f16_to_f32(get16(xq + b * 34));
        for (int k = 0; k < 32; k++) qx[b * 32 + k] = (int8_t)xq[b * 34 + 2 + k];
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < cy; r++) {
        const uint8_t *row = w + r * nb * bb;
        int32_t c[64]; float sc, ba; double acc = 0;
        for (int64_t b = 0; b < nb; b++) {
            orc_block_codes(dt, c, &sc, &ba, row + b * bb);
            for (int h = 0; h < cap / 32; h++) { /* B64 blocks span two activation blocks (gemv.h:1682-1692) */
                int64_t xb = b * (cap / 32) + h;
                long isum = 0, qsum = 0;
                for (int k = 0; k < 32; k++) { isum += (long)c[h * 32 + k] * qx[xb * 32 + k]; qsum += qx[xb * 32 + k]; }
                acc += (double)sx[xb] * ((double)sc * (double)isum + (double)ba * (double)qsum);
            }
        }
        y[r] = acc;
    }
    free(sx); free(qx);
    return 0;
}

int orc_gemv_half(int dt, double *y, const uint8_t *w, const float *x, int64_t cy, int64_t cx) {
    int cap = orc_block_capacity(dt), bb = orc_block_bytes(dt);
    if (cap <= 1) return -1;
    int64_t nb = cx / cap;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < cy; r++) {
        const uint8_t *row = w + r * nb * bb;
        float d[64]; double acc = 0;
        for (int64_t b = 0; b < nb; b++) {
            orc_dequantize_row(dt, d, row + b * bb, cap);
            for (int k = 0; k < cap; k++) acc += (double)(float)(_Float16)d[k] * (double)x[b * cap + k];
        }
        y[r] = acc;
    }
    return 0;
}

int orc_gemv_f16(double *y, const uint16_t *w, const float *x, int64_t cy, int64_t cx) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < cy; r++) {
        double acc = 0;
        for (int64_t k = 0; k < cx; k++) acc += (double)orc_f16_to_f32(w[r * cx + k]) * (double)x[k];
        y[r] = acc;
    }
    return 0;
}

/* ------------------------------------------------------------------ small ops */
void orc_rmsnorm(float *y, const float *x, const float *w, int64_t n, float eps, float multi_base) {
    double s = 0;
    for (int64_t i = 0; i < n; i++) s += (double)x[i] * (double)x[i];
    float mean = (float)(s / (double)n);
    float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = 0; i < n; i++) {
        float v = x[i] * scale;
        if (w) v = v * (multi_base + w[i]);
        y[i] = v;
    }
}

void orc_rope(float *x, int heads, int head_dim, int rope_dims, int pos, float theta, int order) {
    /* theta_scale = theta^(-2/dims); angle_i = pos * theta_scale^i (unary_tensor_opr.h:666-679,706-719) */
    if (rope_dims <= 0) rope_dims = head_dim;
    float theta_scale = powf(theta, -2.0f / (float)rope_dims);
    for (int h = 0; h < heads; h++) {
        float *v = x + (int64_t)h * head_dim;
        for (int i = 0; 2 * i < rope_dims; i++) {
            float ang = (float)pos;
            if (i > 0) ang *= powf(theta_scale, (float)i);
            float c = cosf(ang), s = sinf(ang);
            int i0 = order == 2 ? i : 2 * i;
            int i1 = order == 2 ? i + rope_dims / 2 : 2 * i + 1;
            float x0 = v[i0], x1 = v[i1];
            v[i0] = x0 * c - x1 * s;
            v[i1] = x0 * s + x1 * c;
        }
    }
}

void orc_silu_mul(float *out, const float *a, const float *b, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        float x = a[i];
        float fx = x / (1.0f + expf(-x));
        out[i] = fx * b[i];
    }
}

void orc_attn_decode(float *out, const float *q, const float *k, const float *v, int ctx,
                     int heads, int kv_heads, int head_dim, float kq_scale, int exact) {
    int group = heads / kv_heads;
    int64_t kvd = (int64_t)kv_heads * head_dim;
    float alpha = 1.0f / sqrtf((float)head_dim) / kq_scale; /* inference_worker.cc:1666 */
    double *p = (double *)malloc(sizeof(double) * ctx);
    for (int h = 0; h < heads; h++) {
        int g = h / group;
        const float *qh = q + (int64_t)h * head_dim;
        double mx = -INFINITY;
        for (int t = 0; t < ctx; t++) {
            const float *kr = k + t * kvd + (int64_t)g * head_dim;
            double s = 0;
            for (int d = 0; d < head_dim; d++) s += (double)qh[d] * (double)kr[d];
            s *= alpha;
            if (!exact) s = (float)(_Float16)(float)s; /* kq tensor is F16 */
            s *= kq_scale;                           /* softmax's s_scale (unary_tensor_opr.h:495) */
            p[t] = s;
            if (s > mx) mx = s;
        }
        double sum = 0;
        for (int t = 0; t < ctx; t++) {
            double e = exact ? exp(p[t] - mx) : (double)expf((float)(p[t] - mx));
            sum += e;
            p[t] = exact ? e : (double)(float)(_Float16)(float)e; /* T[offset] = val (half) */
        }
        for (int t = 0; t < ctx; t++) {
            if (exact) p[t] /= sum;
            else { /* T[offset] *= scale : half * half(scale) -> half (unary_tensor_opr.h:533) */
                float sc = (float)(_Float16)(1.0f / (float)sum);
                p[t] = (float)(_Float16)((float)p[t] * sc);
            }
        }
        for (int d = 0; d < head_dim; d++) {
            double acc = 0;
            for (int t = 0; t < ctx; t++) acc += p[t] * (double)v[t * kvd + (int64_t)g * head_dim + d];
            out[(int64_t)h * head_dim + d] = exact ? (float)acc : (float)(_Float16)(float)acc;
        }
    }
    free(p);
}

void orc_moe_topk(int32_t *ids, float *weights, int32_t *cnt, const float *scores, int rows,
                  int experts, int top_k, int norm_top_k_prob, float min_score) {
    if (top_k > 8) top_k = 8; /* RowItemForMoe::MAX_SIZE */
    for (int r = 0; r < rows; r++) {
        const float *s = scores + (int64_t)r * experts;
        int32_t *id = ids + (int64_t)r * top_k; float *wt = weights + (int64_t)r * top_k;
        int n = 0;
        for (int e = 0; e < experts; e++) {
            float sc = s[e];
            if ((n >= top_k && wt[n - 1] >= sc) || sc < min_score) continue;
            if (n < top_k) { id[n] = -1; wt[n] = 0; n++; }
            for (int j = n - 1; j >= 0; j--) { /* insertion, first-wins on ties */
                if (wt[j] >= sc) break;
                if (j + 1 < n) { id[j + 1] = id[j]; wt[j + 1] = wt[j]; }
                id[j] = e; wt[j] = sc;
            }
        }
        cnt[r] = n;
        for (int j = n; j < top_k; j++) { id[j] = -1; wt[j] = 0; }
        if (norm_top_k_prob) {
            float sum = 0;
            for (int j = 0; j < n; j++) sum += wt[j];
            for (int j = 0; j < n; j++) wt[j] /= sum;
        }
    }
}
