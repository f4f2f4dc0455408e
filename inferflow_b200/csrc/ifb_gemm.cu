// Dequant-fused tensor-core GEMM for rows > 1: prefill and batched decode, sm_100a (tcgen05 + TMEM).
//
// Replaces the rows > 1 branch of GpuInferenceWorker::MatrixMultiplication (reference
// src/transformer/inference_worker.cc:2364-2432): Dequantize the whole weight tensor to fp16 (:2375-2380) ->
// cublas_engine_.GemmEx (:2413, src/tensor/cublas_engine.cu:425-505) -> Transpose. Here the weight never exists in
// fp16 in global memory:
//
//   Y[tokens][cy] = X[tokens][cx] . W[cy][cx]^T        (fp16 in, fp32 accumulate in TMEM, fp16 out)
//
//  * a CTA owns 128 * RB weight rows x up to NTMAX tokens (two tile configurations, below); the K loop walks 64 columns
//    at a time through a 3-stage ring of shared-memory tiles;
//  * A tile (weights): the producer warps (thread = one weight row) keep the raw quantized blocks of their row two
//    k-slices ahead in registers, decode them with the same Fmt::decode the GEMV uses, compute
//    value = fp16(q * scale + base) -- the bytes TensorOpr::Dequantize would have written; int -> float through a
//    mantissa trick, the I2F of the first version kept the XU pipe 77 % busy -- and store them straight into the tile;
//  * B tile (activations): ONE 2-D TMA load per k-slice (cuTensorMapEncodeTiled over X, box = 8 halves x NTMAX tokens),
//    issued by a loader warp; both tiles are UMMA "core matrices" (8 rows x 16 bytes, contiguous), K-major, no swizzle;
//  * ONE elected thread of the MMA warp issues tcgen05.mma (M = 128, N = token tile, K = 16 per instruction) with the
//    accumulator in tensor memory; tcgen05.commit hands the shared-memory stage back to the producers;
//  * after the last k-slice the producer warps become the epilogue: tcgen05.ld (32 lanes x 32 bit, lane = weight row)
//    -> fp16 -> Y. The output is written in its final [tokens][cy] orientation (no transpose pass).
//  Measured on B200 (profiles/r2_bench.md, r2_ncu.md): 452-666 TFLOP/s at 4096 tokens, tensor pipe active 20-22 % of
//  the cycles -- the kernel is bound by producing fp16 operands from the 4-bit blocks (ALU pipe 18-43 %).
#include <map>
#include <mutex>
#include <utility>

#include <cuda.h>      // CUtensorMap + the driver's encode function (fetched through cudaGetDriverEntryPoint: no libcuda link)

#include "ifb_formats.cuh"
#include "ifb_kernels.h"

namespace ifb {

constexpr int GQ_ROWS = 128;        // weight rows per accumulator = UMMA M
constexpr int GQ_BK = 64;           // k-slice per stage
constexpr int GQ_STAGES = 3;
constexpr int GQ_A_BYTES = GQ_ROWS * GQ_BK * 2;      // 16 KB per 128-row block
constexpr int GQ_STRIP = 80;        // per-thread scratch for the raw blocks of one k-slice (<= 2 x 36 B + slack)
// Two tile configurations (template parameters RB = 128-row blocks per CTA, NT = token tile):
//   small (RB 1, NT <= 128): 128 rows x 128 tokens, 4 producer warps, 96 KB of shared memory -> two CTAs per SM hide
//         each other's fetch latency; for few tokens (batched decode), where the weight stream is the cost;
//   large (RB 2, NT 256):   256 rows x 256 tokens, 8 producer warps, both accumulators fill the 512 TMEM columns; a
//         weight element is dequantized once per 256 tokens (the dequantization is ALU work that competes with nothing
//         the tensor core needs, but at 128 tokens per pass it was what bound the first version: ~190 TFLOP/s).

// ---- tcgen05 / TMEM wrappers (PTX; SASS: UTCHMMA, LDTM, UTCBAR ...)
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier receives one arrival when every tcgen05 operation this thread issued so far has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 bit x 16 columns: thread (lane) i of the warp gets columns c..c+15 of TMEM lane (warp%4)*32 + i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 2-D TMA tile load (SASS UTMALDG): box of the tensor map at (c0, c1) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// unsigned byte -> float without the conversion unit: drop the byte into the mantissa of 2^23, subtract 2^23 (exact).
// (ncu on the first version: the XU pipe, where I2F lives, was 77 % busy -- one I2F per weight element.)
__device__ __forceinline__ float u8f(uint32_t word, int k) {
    return __uint_as_float(__byte_perm(word, 0x4b000000u, 0x7650u + (uint32_t)k)) - 8388608.0f;
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (K-adjacent core matrices), [32,46) stride byte
// offset >> 4 (8-row groups), [46,48) version = 1, [61,64) layout type 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

struct GemmArgs {
    const uint8_t *w;        // [cy][cx / cap] blocks (or fp16 [cy][cx])
    const __half *x;         // [tokens][cx]
    __half *y;               // [tokens][cy]
    const __half *bias;      // [cy] or null
    int tokens, cy, cx, row_bytes, w_dtype;
    int nt;                  // token tile (multiple of 16)
    // split-K (few tokens: cy / 128 CTAs would leave most SMs idle while 128 threads per CTA dequantize whole rows):
    // blockIdx.z takes nk_per k-slices and leaves fp32 partial sums in part[z][tokens][cy]; gemm_splitk_reduce_kernel
    // adds them in split order (deterministic) and applies the half rounding + bias of the one-pass epilogue
    int nk_per;              // k-slices per split (0: no split)
    float *part;
};

// element (r, k) of a tile: core matrices of 8 rows x 8 halves (128 contiguous bytes); K-adjacent core matrices 128 B
// apart (LBO), 8-row groups (BK/8) * 128 B apart (SBO)
constexpr int GQ_LBO = 128, GQ_SBO = (GQ_BK / 8) * 128;

// B tile (activations), written by TMA: chunk k8 (8 halves of every token) is one [NTMAX][16 B] box, boxes 16*NTMAX apart:
//   element (t, k) at (k / 8) * 16 * NTMAX + t * 16 + (k % 8) * 2   ->   LBO = 16 * NTMAX, SBO = 128 (8 tokens x 16 B)
template <class Fmt, bool F16W, int RB, int NTMAX>
__global__ void __launch_bounds__(128 * RB + 64, RB == 1 ? 2 : 1) gemm_q_kernel(const GemmArgs a, const __grid_constant__ CUtensorMap xmap) {
    constexpr int NPROD = 128 * RB;                       // producer threads: thread = one weight row of the CTA's rows
    constexpr int PW = NPROD / 32;                        // producer warps; warp PW issues the MMAs, warp PW+1 the TMA loads of X
    constexpr int B_BYTES = NTMAX * GQ_BK * 2;
    constexpr int STAGE_BYTES = RB * GQ_A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = RB * NTMAX < 32 ? 32 : RB * NTMAX;
    constexpr uint32_t B_LBO = 16 * NTMAX, B_SBO = 128;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[GQ_STAGES], empty_bar[GQ_STAGES], done_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * (GQ_ROWS * RB), tok0 = blockIdx.y * a.nt;
    const int nk_all = a.cx / GQ_BK;
    const int kb0 = a.nk_per > 0 ? (int)blockIdx.z * a.nk_per : 0;            // first k-slice of this CTA
    const int nk = a.nk_per > 0 ? min(a.nk_per, nk_all - kb0) : nk_all;
    uint8_t *strips = smem + GQ_STAGES * STAGE_BYTES;     // [NPROD][GQ_STRIP]

    if (tid == 0) {
        // a stage is full when every producer warp has delivered its rows of A (PW arrivals) and the TMA bytes of B landed
        for (int i = 0; i < GQ_STAGES; i++) { mbar_init(&full_bar[i], PW + 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == PW) tmem_alloc(&tmem_base_s, TMEM_COLS);  // fp32 accumulators: RB x nt columns (power of two >= 32)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == PW + 1) {
        // ===== activation loader: one thread, runs up to GQ_STAGES slices ahead =====
        if (lane == 0) {
            for (int kb = 0; kb < nk; kb++) {
                const int s = kb % GQ_STAGES;
                mbar_wait(&empty_bar[s], ((uint32_t)(kb / GQ_STAGES) & 1u) ^ 1u);
                uint8_t *Bt = smem + s * STAGE_BYTES + RB * GQ_A_BYTES;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)B_BYTES);
#pragma unroll
                for (int k8 = 0; k8 < GQ_BK / 8; k8++)           // tokens beyond the input are zero-filled by the TMA unit
                    tma_load_2d(Bt + k8 * B_LBO, &xmap, (kb0 + kb) * GQ_BK + k8 * 8, tok0, &full_bar[s]);
            }
        }
    } else if (warp == PW) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            // instruction descriptor (InstrDescriptor): c_format F32 (bits 4-5 = 1), a/b fp16 K-major (0), N >> 3 at bit 17,
            // M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.nt >> 3) << 17) | ((uint32_t)(GQ_ROWS >> 4) << 24);
            for (int kb = 0; kb < nk; kb++) {
                const int s = kb % GQ_STAGES;
                mbar_wait(&full_bar[s], (uint32_t)(kb / GQ_STAGES) & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + RB * GQ_A_BYTES;
#pragma unroll
                for (int k = 0; k < GQ_BK / 16; k++) {
                    // one instruction consumes K = 16: two K-adjacent core matrices = 2 * LBO bytes further
                    const uint64_t bd = umma_desc(b_addr + k * 2 * B_LBO, B_LBO, B_SBO);
#pragma unroll
                    for (int rb = 0; rb < RB; rb++) {
                        const uint64_t ad = umma_desc(a_addr + rb * GQ_A_BYTES + k * 2 * GQ_LBO, GQ_LBO, GQ_SBO);
                        umma_f16(tmem + (uint32_t)(rb * a.nt), ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);          // the stage is free once these MMAs have read it
            }
            umma_commit(&done_bar);                  // accumulators complete
        }
    } else {
        // ===== producers: thread = weight row `tid` of the CTA's rows =====
        const int r = tid, rb = r >> 7, rr = r & 127;
        const uint8_t *wrow = a.w + (size_t)(row0 + r) * a.row_bytes;
        uint8_t *strip = strips + r * GQ_STRIP;
        constexpr int NBLK = F16W ? 1 : GQ_BK / Fmt::CAP, SEG = F16W ? GQ_BK * 2 : NBLK * Fmt::BYTES;
        constexpr int NRAW = F16W ? GQ_BK * 2 / 4 : (SEG + 3 + 3) / 4;     // 32-bit words fetched per k-slice
        // the row's blocks of a k-slice are fetched into registers TWO slices ahead: the L2 / HBM round trip (the
        // first version stalled ~3 us per slice on it) runs under two slices of dequantization and MMAs
        uint32_t raw0[NRAW], raw1[NRAW], raw2[NRAW];
        auto fetch = [&](int kb, uint32_t *rw) {
            if (kb >= nk) return;
            const int k0 = (kb0 + kb) * GQ_BK;
            if (F16W) {
                const uint4 *src = reinterpret_cast<const uint4 *>(wrow + (size_t)k0 * 2);
#pragma unroll
                for (int i = 0; i < GQ_BK / 8; i++) {
                    const uint4 v = src[i];
                    rw[4 * i] = v.x; rw[4 * i + 1] = v.y; rw[4 * i + 2] = v.z; rw[4 * i + 3] = v.w;
                }
            } else {
                const uint8_t *seg = wrow + (size_t)(k0 / Fmt::CAP) * Fmt::BYTES;
                const uint32_t m = (uint32_t)((uintptr_t)seg & 3u);       // blocks are only 2- or 4-byte aligned
                const uint32_t *seg4 = reinterpret_cast<const uint32_t *>(seg - m);
#pragma unroll
                for (int i = 0; i < NRAW; i++) rw[i] = (i * 4 < (int)m + SEG) ? seg4[i] : 0u;   // never a word past the row's bytes
            }
        };
        auto produce = [&](int kb, const uint32_t *raw) {
            const int s = kb % GQ_STAGES;
            mbar_wait(&empty_bar[s], ((uint32_t)(kb / GQ_STAGES) & 1u) ^ 1u);
            uint8_t *At = smem + s * STAGE_BYTES + rb * GQ_A_BYTES;
            uint8_t *arow = At + (rr >> 3) * GQ_SBO + (rr & 7) * 16;         // + k8 * LBO
            if (F16W) {
#pragma unroll
                for (int k8 = 0; k8 < GQ_BK / 8; k8++)
                    *reinterpret_cast<uint4 *>(arow + k8 * GQ_LBO) = make_uint4(raw[4 * k8], raw[4 * k8 + 1], raw[4 * k8 + 2], raw[4 * k8 + 3]);
            } else {
                // raw bytes -> this thread's strip, decoded with the GEMV's Fmt::decode
                const uint32_t mis = (uint32_t)((uintptr_t)(wrow + (size_t)((kb0 + kb) * GQ_BK / Fmt::CAP) * Fmt::BYTES) & 3u);
                uint32_t *strip4 = reinterpret_cast<uint32_t *>(strip);
#pragma unroll
                for (int i = 0; i < NRAW; i++) strip4[i] = raw[i];
                __syncwarp();
#pragma unroll
                for (int b = 0; b < NBLK; b++) {
                    uint32_t codes[Fmt::NW];
                    float sc, ba;
                    Fmt::decode(strip + mis + b * Fmt::BYTES, codes, sc, ba);
#pragma unroll
                    for (int c8 = 0; c8 < Fmt::CAP / 8; c8++) {          // 8 elements = one 16-byte chunk of the A tile
                        uint4 out;
                        __half2 *oh = reinterpret_cast<__half2 *>(&out);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            // element e of the block sits in byte e%4 of codes[e/4] ... except for the nibble-split formats,
                            // whose staged order is (x0,x2,x4,x6)(x1,x3,x5,x7) per group of 8 (see x_pos in ifb_gemv_dev.cuh)
                            float q0, q1;
                            if (Fmt::XSPLIT) {
                                q0 = u8f(codes[2 * c8], j);                  // element 8*c8 + 2j
                                q1 = u8f(codes[2 * c8 + 1], j);              // element 8*c8 + 2j + 1
                            } else if (Fmt::SIGNED) {
                                const int e = 8 * c8 + 2 * j;                // signed codes: bias by 128, take it back exactly
                                q0 = u8f(codes[e >> 2] ^ 0x80808080u, e & 3) - 128.0f;
                                q1 = u8f(codes[(e + 1) >> 2] ^ 0x80808080u, (e + 1) & 3) - 128.0f;
                            } else {
                                const int e = 8 * c8 + 2 * j;
                                q0 = u8f(codes[e >> 2], e & 3);
                                q1 = u8f(codes[(e + 1) >> 2], (e + 1) & 3);
                            }
                            // TensorOpr::Dequantize: value = q * scale + base in fp32, stored as half
                            oh[j] = __floats2half2_rn(fmaf(q0, sc, ba), fmaf(q1, sc, ba));
                        }
                        *reinterpret_cast<uint4 *>(arow + (b * (Fmt::CAP / 8) + c8) * GQ_LBO) = out;
                    }
                }
            }
            fence_async_smem();                      // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        };
        fetch(0, raw0);
        fetch(1, raw1);
        for (int kb = 0; kb < nk; kb += 3) {         // register sets rotate by name: the loop is unrolled by three
            fetch(kb + 2, raw2);
            produce(kb, raw0);
            if (kb + 1 < nk) { fetch(kb + 3, raw0); produce(kb + 1, raw1); }
            if (kb + 2 < nk) { fetch(kb + 4, raw1); produce(kb + 2, raw2); }
        }
        // ===== epilogue: TMEM lane = weight row; warp w reads lane quadrant w % 4 of accumulator w / 4 =====
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        const float bias = a.bias != nullptr ? __half2float(a.bias[row0 + r]) : 0.f;
        for (int c0 = 0; c0 < a.nt; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(rb * a.nt + c0), v);
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const int t = tok0 + c0 + c;
                if (t < a.tokens) {
                    float f = __uint_as_float(v[c]);
                    if (a.part != nullptr) {
                        a.part[((size_t)blockIdx.z * a.tokens + t) * a.cy + row0 + r] = f;
                        continue;
                    }
                    if (a.bias != nullptr) f = __half2float(__float2half_rn(f)) + bias;     // GEMM output is half, then + bias
                    a.y[(size_t)t * a.cy + row0 + r] = __float2half_rn(f);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == PW) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

__global__ void gemm_splitk_reduce_kernel(__half *y, const float *part, const __half *bias, int tokens, int cy, int ksplit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)tokens * cy) return;
    float f = 0.f;
    for (int s = 0; s < ksplit; s++) f += part[(size_t)s * tokens * cy + i];      // split order: the same sum every run
    if (bias != nullptr) f = __half2float(__float2half_rn(f)) + __half2float(bias[i % cy]);
    y[i] = __float2half_rn(f);
}

// the activation matrix [tokens][cx] fp16 as a 2-D tensor map with a box of NTMAX tokens x 8 halves (16 bytes): one
// TMA load lands chunk k8 of every token of the tile contiguously -- the UMMA core-matrix layout of the B operand
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_xmap(CUtensorMap *map, const void *x, int64_t tokens, int64_t cx, int ntmax) {
    static EncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        IFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        IFB_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "gemm: the driver does not export cuTensorMapEncodeTiled");
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cx, (cuuint64_t)tokens};
    const cuuint64_t strides[1] = {(cuuint64_t)cx * 2};
    const cuuint32_t box[2] = {8, (cuuint32_t)ntmax};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("gemm: cuTensorMapEncodeTiled failed (%d) for x [%lld][%lld]", (int)r, (long long)tokens, (long long)cx);
        return IFB_ERR_CUDA;
    }
    return IFB_OK;
}

static int splitk_scratch(cudaStream_t st, size_t need, float **out) {
    struct Buf { void *p; size_t bytes; };
    static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Buf &b = bufs[std::make_pair(dev, st)];
    if (b.bytes < need) {
        if (b.p != nullptr) { cudaStreamSynchronize(st); cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
        const size_t bytes = (need + ((size_t)1 << 20) - 1) >> 20 << 20;
        IFB_CUDA(cudaMalloc(&b.p, bytes));
        b.bytes = bytes;
    }
    *out = (float *)b.p;
    return IFB_OK;
}

template <class Fmt, bool F16W, int RB, int NTMAX>
static int launch_gemm_cfg(const GemmArgs &a, cudaStream_t st) {
    auto kern = gemm_q_kernel<Fmt, F16W, RB, NTMAX>;
    const size_t smem = (size_t)GQ_STAGES * (RB * GQ_A_BYTES + NTMAX * GQ_BK * 2) + (size_t)128 * RB * GQ_STRIP;
    static int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev = dev;
    }
    CUtensorMap xmap;
    int rc = make_xmap(&xmap, a.x, a.tokens, a.cx, NTMAX);
    if (rc != IFB_OK) return rc;
    dim3 grid(a.cy / (GQ_ROWS * RB), (a.tokens + a.nt - 1) / a.nt);
    // few tokens: split K until ~two CTAs per SM are busy (each split >= 4 k-slices, at most 8 splits)
    const int ctas = (int)(grid.x * grid.y), nk_all = a.cx / GQ_BK;
    int ksplit = 1;
    if (RB == 1 && a.tokens <= 32 && ctas < 2 * 148) {
        ksplit = (4 * 148 + ctas - 1) / ctas;
        if (ksplit > 8) ksplit = 8;
        if (ksplit > nk_all / 4) ksplit = nk_all / 4;
        if (ksplit < 1) ksplit = 1;
    }
    if (ksplit > 1) {
        GemmArgs b = a;
        b.nk_per = (nk_all + ksplit - 1) / ksplit;
        ksplit = (nk_all + b.nk_per - 1) / b.nk_per;
        // scratch for the partial sums: one buffer per (device, stream), grown when a larger product comes (the first
        // version asked the stream-ordered allocator per call: with its default release threshold every call paid a
        // real cudaMalloc -- 0.5-1.4 ms against 25 us of work)
        const size_t need = (size_t)ksplit * a.tokens * a.cy * sizeof(float);
        rc = splitk_scratch(st, need, &b.part);
        if (rc != IFB_OK) return rc;
        grid.z = ksplit;
        kern<<<grid, 128 * RB + 64, smem, st>>>(b, xmap);
        const int64_t n = (int64_t)a.tokens * a.cy;
        gemm_splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.y, b.part, a.bias, a.tokens, a.cy, ksplit);
        IFB_LAUNCH_CHECK();
        return IFB_OK;
    }
    kern<<<grid, 128 * RB + 64, smem, st>>>(a, xmap);
    IFB_LAUNCH_CHECK();
    return IFB_OK;
}

template <class Fmt, bool F16W>
static int launch_gemm(GemmArgs a, cudaStream_t st) {
    const int t16 = (a.tokens + 15) / 16 * 16;
    if (a.tokens > 128 && a.cy % 256 == 0) {
        a.nt = t16 > 256 ? 256 : t16;
        return launch_gemm_cfg<Fmt, F16W, 2, 256>(a, st);
    }
    a.nt = t16 > 128 ? 128 : t16;
    return launch_gemm_cfg<Fmt, F16W, 1, 128>(a, st);
}

// Y[tokens][cy] = X[tokens][cx] . W[cy][cx]^T (+ bias). W: any dp4a-capable block format (or fp16).
int gemm_q(int w_dtype, void *y_f16, const void *w, const void *x_f16, int64_t tokens, int64_t cy, int64_t cx,
           const void *bias_f16, cudaStream_t st) {
    IFB_CHECK_ARG(tokens >= 1 && cy >= 1 && cx >= 1, "gemm: empty shape");
    IFB_CHECK_ARG(cy % GQ_ROWS == 0, "gemm: rows of W (%lld) must be a multiple of %d", (long long)cy, GQ_ROWS);
    IFB_CHECK_ARG(cx % GQ_BK == 0, "gemm: cx (%lld) must be a multiple of %d", (long long)cx, GQ_BK);
    IFB_CHECK_ARG((((uintptr_t)x_f16 | (uintptr_t)w) & 15) == 0 || w_dtype != DT_F16, "gemm: fp16 operands must be 16-byte aligned");
    IFB_CHECK_ARG(((uintptr_t)x_f16 & 15) == 0, "gemm: x must be 16-byte aligned");
    GemmArgs a;
    a.w = (const uint8_t *)w; a.x = (const __half *)x_f16; a.y = (__half *)y_f16; a.bias = (const __half *)bias_f16;
    a.tokens = (int)tokens; a.cy = (int)cy; a.cx = (int)cx; a.w_dtype = w_dtype;
    a.row_bytes = (int)(w_dtype == DT_F16 ? cx * 2 : row_bytes(w_dtype, cx));
    a.nt = 0; a.nk_per = 0; a.part = nullptr;
    switch (w_dtype) {
        case DT_F16: return launch_gemm<FmtQ4_B32T1, true>(a, st);
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: return launch_gemm<FmtQ4_B32T1, false>(a, st);
        case DT_Q4_B64T1: return launch_gemm<FmtQ4_B64T1, false>(a, st);
        case DT_Q8_B32T2: return launch_gemm<FmtQ8_B32T2, false>(a, st);
        case DT_Q6_B64T1: return launch_gemm<FmtQ6_B64T1, false>(a, st);
        case DT_Q5_B64T1: return launch_gemm<FmtQ5_B64T1, false>(a, st);
        case DT_Q3H_B64T1: return launch_gemm<FmtQ3H_B64T1, false>(a, st);
        default:
            set_last_error("gemm: weight type %d has no tensor-core path (reference: Dequantize + cublasGemmEx)", w_dtype);
            return IFB_ERR_TYPE;
    }
}

}  // namespace ifb
