// Decode GEMV for Inferflow block-quantized weights, sm_100a.
//
// Replaces TensorMul::Gemv_AX -> Gemv_AX_Alg3 -> Gemv_AX_QuantQ8 / Gemv_AX_QuantHalf / fp16
// (reference src/tensor/tensor_mul.cu:97,766,714,658; kernels src/kernels/gemv.h:470-1709)
// and, in its fused form, the RmsNorm -> Quantize(Q8_B32T2) -> GEMV(s) -> SiLU*Mul / +residual
// chains of GpuInferenceWorker (src/transformer/inference_worker.cc:1038,1442-1444,
// 1555-1571,1726-1900).
//
// Design (HBM-bound streaming kernel; DESIGN.md "gemv"):
//  * ONE persistent CTA per SM whose weight ring takes (almost) the whole SM's shared memory (216 KB; measured: bytes
//    in flight per SM matter more than a co-resident successor -- 110 KB 512 tok/s, 150 KB 559, 216 KB 586). Only the
//    wq|wk|wv launch keeps 150 KB so that the attention kernel's CTAs fit beside it and prefetch the KV cache early.
//    Under programmatic dependent launch the NEXT kernel's producer warp fills its ring as soon as an SM has room;
//  * 1 producer warp + 16 consumer warps. The weight matrix is streamed in the reference's own
//    AoS block layout: a "unit" is R whole rows = one contiguous byte range, moved
//    global->shared by ONE 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) into a ring of
//    shared-memory slots guarded by full/empty mbarriers; units are handed out by a global
//    cursor (SMs do not get equal shares of the memory bandwidth);
//  * consumers: one warp per row; for the headline format (Q4_B32T1, 20-byte blocks) a lane owns 4
//    consecutive blocks = 80 bytes = five conflict-free LDS.128; integer dot by dp4a against the
//    int8 activation staged once per CTA (high nibbles are used in place as 16*q: no shift),
//    fp32 combine, warp-shuffle reduce;
//  * the activation prologue (optional RMSNorm, Q8_B32T2 quantization identical to
//    Tensor_QuantizeQ8_B32T2_Alg2_Kernel, kernels/tensor_quant.h:45-82) runs in every CTA
//    from the fp16 vector with one memory round trip.
#include "ifb_formats.cuh"
#include "ifb_gemv_dev.cuh"
#include "ifb_kernels.h"
#include <cstdlib>

namespace ifb {

constexpr int GEMV_THREADS = GEMV_CTHREADS + 32;      // consumer warps (ifb_gemv_dev.cuh) + producer warp
constexpr int GEMV_MAX_SLOTS = 16;

struct GemvArgs {
    const uint8_t *w0, *w1, *w2;   // up to 3 segments sharing x
    int unit1, unit2;              // first unit of segment 1 / 2 (INT_MAX when absent)
    long long bytes0, bytes1, bytes2;   // bytes of each segment (L2 prefetch)
    const uint8_t *w_gate;
    int cx, nblk, row_bytes;
    int nxb_alloc;   // activation blocks allocated in shared memory (padded to 128 for the transposed layout)
    int R;           // rows per unit/slot
    int wpg;         // consumer warps per group (= min(R, CWARPS))
    int groups;      // consumer groups in use; nslots % groups == 0 so a group meets every fill of its slots in order
    int nslots, slot_stride;
    int units;
    int l2_prefetch; // >0: software prefetch distance in UNITS: unit u+dist is requested into the L2 when unit u is copied
    int xg_off;      // byte offset (from the staged-x area) of the fp16 vector a tensor-parallel gather builds
    int pdl_late;
    int batch_claim;
    int debug_nocompute;  // measurement aid (IFB_GEMV_NOCOMPUTE=1): consumers release slots without doing the math
    const __half *x;
    const uint8_t *xq;  // pre-quantized activation (Q8_B32T2 blocks) or null
    const __half *norm_w;
    float eps, norm_base;
    const __half *residual, *bias;
    __half *y;
    int w_dtype;
    unsigned long long *trace;
    int *counter;    // work-stealing cursor over the units (zero before the launch)
    TpExchange tp;   // tensor-parallel push / gather (size 1: unused)
    // mixture of experts: the matrices of this launch are picked on the device (see GemvLaunch)
    const uint8_t *const *w_table, *const *wg_table;
    const int32_t *moe_id;
};

// Tensor-parallel gather of one group of 8 activations: residual + partial sums of all ranks, added in
// rank order with fp16 rounding after every add (MergeTensors adds worker 1..n-1 onto worker 0 with
// TensorOpr::Add, then the layer adds the residual: inference_worker.cc:2264-2289, :849).
// The partials arrive as flag-in-data words {half2, flag} written by the peers' push epilogues: a thread polls
// exactly the 4 words (32 bytes) per rank it needs -- no fence, no CTA-wide wait.
__device__ __forceinline__ uint32_t tp_flag(const int32_t *seq, uint32_t idx) {
    return ((uint32_t)*reinterpret_cast<const volatile int32_t *>(seq) << 8) | idx;
}
__device__ __forceinline__ uint4 tp_gather8(const GemvArgs &a, int g, uint32_t flag) {
    const uint8_t *base = reinterpret_cast<const uint8_t *>(a.tp.peer[a.tp.rank]) + a.tp.src_off;
    float acc[8];
#pragma unroll
    for (int p0 = 0; p0 < 8; p0 += 4) {
        if (p0 >= a.tp.size) break;
        uint4 w[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (p0 + i < a.tp.size) {
                const uint4 *src = reinterpret_cast<const uint4 *>(base + (size_t)(p0 + i) * a.tp.src_rank_stride) + 2 * g;
                w[i][0] = ld_relaxed_sys_v4(src);
                w[i][1] = ld_relaxed_sys_v4(src + 1);
            }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (p0 + i < a.tp.size) {
                const uint4 *src = reinterpret_cast<const uint4 *>(base + (size_t)(p0 + i) * a.tp.src_rank_stride) + 2 * g;
                PeerDeadline dl;   // bounded: see ifb_common.cuh
                while (w[i][0].y != flag || w[i][0].w != flag) { w[i][0] = ld_relaxed_sys_v4(src); dl.tick(); }
                while (w[i][1].y != flag || w[i][1].w != flag) { w[i][1] = ld_relaxed_sys_v4(src + 1); dl.tick(); }
                const uint32_t d[4] = {w[i][0].x, w[i][0].z, w[i][1].x, w[i][1].z};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&d[k]));
                    if (p0 + i == 0) { acc[2 * k] = f.x; acc[2 * k + 1] = f.y; }
                    else { acc[2 * k] = round_h(acc[2 * k] + f.x); acc[2 * k + 1] = round_h(acc[2 * k + 1] + f.y); }
                }
            }
    }
    const uint4 rv = reinterpret_cast<const uint4 *>(a.tp.resid)[g];
    uint4 out;
    __half2 *o = reinterpret_cast<__half2 *>(&out);
    const __half2 *r = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 rf = __half22float2(r[i]);
        o[i] = __floats2half2_rn(acc[2 * i] + rf.x, acc[2 * i + 1] + rf.y);
    }
    if (blockIdx.x == 0) reinterpret_cast<uint4 *>(a.tp.xout)[g] = out;
    return out;
}


// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
enum { MODE_Q8 = 0, MODE_HX = 1, MODE_F16 = 2 };

template <class Fmt, int MODE, bool GLU, bool VEC4, bool TP>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_stream_kernel(const GemvArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[GEMV_MAX_SLOTS], empty_bar[GEMV_MAX_SLOTS];
    __shared__ float red[GEMV_CWARPS];
    __shared__ int slot_unit[GEMV_MAX_SLOTS];   // which unit sits in each ring slot (-1: no more work)
    __shared__ __align__(16) __half push_y[GEMV_MAX_SLOTS][32];   // tensor-parallel push: a unit's R outputs, per ring slot
    __shared__ unsigned int push_cnt[GEMV_MAX_SLOTS];             // ... and how many of the group's warps have delivered theirs

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *ring = smem;
    uint8_t *xbase = smem + (size_t)a.nslots * a.slot_stride;
    const int nxb = a.nxb_alloc;
    // MODE_Q8: xlo | xhi | xs ; else: xh (fp16)
    uint8_t *xlo = xbase, *xhi = xbase + (size_t)nxb * 16;
    float2 *xs = reinterpret_cast<float2 *>(xbase + (size_t)nxb * 32);
    __half *xh = reinterpret_cast<__half *>(xbase);

    if (tid == 0) {
        for (int i = 0; i < a.nslots; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], a.wpg);
            push_cnt[i] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();
    // PDL: let the next kernel's CTAs become resident as early as possible -- their producer
    // warps only touch weights, everything else sits behind their own griddepcontrol.wait.
    if (!a.pdl_late) pdl_launch_dependents();
    if (tid == 0) trace_mark(a.trace, 0);

    const int nmat = GLU ? 2 : 1;
    const uint32_t mat_bytes = (uint32_t)a.R * (uint32_t)a.row_bytes;

    if (warp == GEMV_CWARPS) {
        // ===== producer: independent of the previous kernel =====
        // Software prefetch HBM -> L2, a fixed distance ahead of the copies into the ring: the bulk
        // copies then hit the L2 (short latency: the shallow ring is enough to keep the stream full)
        // and the DRAM traffic stays in consumption order instead of racing the running kernel.
        const uint8_t *w0p = a.w0, *wgp = a.w_gate;
        if (a.w_table != nullptr) {
            // expert weights: the id is written by the routing kernel earlier in this step, so this producer cannot
            // run ahead of the grid dependency
            pdl_wait();
            const int e = *a.moe_id;
            if (e < 0) return;
            w0p = a.w_table[e];
            if (a.wg_table != nullptr) wgp = a.wg_table[e];
        }
        auto unit_src = [&](int u) -> const uint8_t * {
            if (u >= a.unit2) return a.w2 + (size_t)(u - a.unit2) * mat_bytes;
            if (u >= a.unit1) return a.w1 + (size_t)(u - a.unit1) * mat_bytes;
            return w0p + (size_t)u * mat_bytes;
        };
        if (a.l2_prefetch > 0) {
            // the first `dist` units: shared out across CTAs and lanes, issued at kernel entry
            for (int u = blockIdx.x * 32 + lane; u < a.l2_prefetch && u < a.units; u += gridDim.x * 32) {
                bulk_prefetch_l2(unit_src(u), mat_bytes);
                if (GLU) bulk_prefetch_l2(wgp + (size_t)u * mat_bytes, mat_bytes);
            }
        }
        // Units are handed out by a global cursor (atomicAdd): SMs do not get equal shares of the
        // memory bandwidth, with a static split the slowest SM sets the kernel time.
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            int pos = 0;
            uint32_t phase = 1;   // a fresh mbarrier passes a wait on parity 1
            // The cursor is a global atomic (~0.6 us round trip): claim a whole ring's worth of units at once so the
            // empty ring fills without waiting for one atomic per slot, then one unit at a time, fetched while
            // the producer waits for a free slot.
            const int nb = a.batch_claim ? a.nslots : 1;
            int batch = atomicAdd(a.counter, nb), left = nb, next = -1;
            int sentinels = 0;
            for (;;) {
                int u;
                if (left > 0) { u = batch++; left--; }
                else { if (next < 0) next = atomicAdd(a.counter, 1); u = next; next = -1; }
                const bool real = u < a.units;
                if (real && left == 0) next = atomicAdd(a.counter, 1);   // in flight while we wait for the slot
                mbar_wait(&empty_bar[pos], phase);
                slot_unit[pos] = real ? u : -1;
                if (real) {
                    const uint8_t *src = unit_src(u);
                    uint8_t *dst = ring + (size_t)pos * a.slot_stride;
                    if (a.l2_prefetch > 0 && u + a.l2_prefetch < a.units) {
                        bulk_prefetch_l2(unit_src(u + a.l2_prefetch), mat_bytes);
                        if (GLU) bulk_prefetch_l2(wgp + (size_t)(u + a.l2_prefetch) * mat_bytes, mat_bytes);
                    }
                    mbar_arrive_expect_tx(&full_bar[pos], mat_bytes * nmat);
                    bulk_g2s_hint(dst, src, mat_bytes, &full_bar[pos], pol);
                    if (GLU)
                        bulk_g2s_hint(dst + mat_bytes, wgp + (size_t)u * mat_bytes, mat_bytes,
                                      &full_bar[pos], pol);
                } else {
                    mbar_arrive(&full_bar[pos]);   // an empty slot: tells one consumer group to stop
                    if (++sentinels == a.groups) break;
                }
                if (++pos == a.nslots) { pos = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ===== consumers =====
    pdl_wait();  // x / residual come from the previous kernel
    if (tid == 0) trace_mark(a.trace, 1);
    if (a.w_table != nullptr && *a.moe_id < 0) return;   // unused expert slot of this token
    uint32_t pflag = 0u;
    const void *xsrc = a.x;
    if constexpr (TP) {
        if (a.tp.push) pflag = tp_flag(a.tp.seq, a.tp.push_idx);
        if (a.tp.gather) {
            // tensor-parallel input: x = residual + sum of the ranks' partial vectors, built once into shared
            // memory (fp16); the staging below then reads it from there exactly as it reads a.x otherwise
            uint4 *xg = reinterpret_cast<uint4 *>(xbase + a.xg_off);
            const uint32_t gflag = tp_flag(a.tp.seq, a.tp.gather_idx);
            for (int g = tid; g < (a.cx >> 3); g += GEMV_CTHREADS) xg[g] = tp_gather8(a, g, gflag);
            named_bar_sync(BAR_CONSUMERS, GEMV_CTHREADS);
            xsrc = xg;
        }
    }
    const XStage xst = {a.cx, a.xq, a.norm_w, a.eps, a.norm_base};
    if (MODE == MODE_Q8)
        stage_x_q8<Fmt::XSPLIT, VEC4>(xst, xsrc, xlo, xhi, xs, red, tid);
    else
        stage_x_f16(xst, xsrc, xh, red, tid);
    named_bar_sync(BAR_CONSUMERS, GEMV_CTHREADS);
    if (a.pdl_late) pdl_launch_dependents();   // experiment (IFB_PDL_LATE=1): dependents start after our staging
    if (tid == 0) trace_mark(a.trace, 2);

    const int groups = a.groups;
    const int grp = warp / a.wpg, wl = warp - grp * a.wpg;
    if (grp >= groups) return;   // spare warps (ring too shallow for every group)
    const int cap = block_capacity(a.w_dtype), bb = block_bytes(a.w_dtype);

    // group g consumes the CTA's slot sequence j = g, g+groups, ...; ring position / phase advance with j
    int pos = grp % a.nslots;
    uint32_t phase = (uint32_t)(grp / a.nslots) & 1u;
    const int step_pos = groups % a.nslots, step_wrap = groups / a.nslots;
    for (int j = grp;; j += groups) {
        mbar_wait(&full_bar[pos], phase);
        const int u = slot_unit[pos];
        if (u < 0) break;
        if (tid == 0 && j == 0) trace_mark(a.trace, 3);
        const uint8_t *slot = ring + (size_t)pos * a.slot_stride;
        const int out_base = u * a.R;   // segments are concatenated in y and are multiples of R

        for (int r = wl; r < a.R && !a.debug_nocompute; r += a.wpg) {
            const uint8_t *row = slot + (size_t)r * a.row_bytes;
            const int o = out_base + r;
            // lane 0 fetches the epilogue operands now and converts them after the reduction
            unsigned short res_bits = 0, bias_bits = 0;
            if (lane == 0) {
                if (a.residual != nullptr) res_bits = reinterpret_cast<const unsigned short *>(a.residual)[o];
                if (a.bias != nullptr) bias_bits = reinterpret_cast<const unsigned short *>(a.bias)[o];
            }
            float acc[2] = {0.f, 0.f};
            if (MODE == MODE_Q8) {
                if (VEC4) {
                    row_dot_q4b32_vec4<GLU ? 2 : 1>(row, mat_bytes, a.nblk, (const uint4 *)xlo,
                                                    (const uint4 *)xhi, xs, lane, acc);
                } else {
                    acc[0] = row_dot_q8<Fmt>(row, a.nblk, (const uint4 *)xlo, (const uint4 *)xhi, xs, lane);
                    if (GLU) acc[1] = row_dot_q8<Fmt>(row + mat_bytes, a.nblk, (const uint4 *)xlo,
                                                      (const uint4 *)xhi, xs, lane);
                }
            } else if (MODE == MODE_HX) {
                acc[0] = row_dot_hx(a.w_dtype, row, a.nblk, cap, bb, xh, lane);
                if (GLU) acc[1] = row_dot_hx(a.w_dtype, row + mat_bytes, a.nblk, cap, bb, xh, lane);
            } else {
                acc[0] = row_dot_f16(row, a.cx, xh, lane);
                if (GLU) acc[1] = row_dot_f16(row + mat_bytes, a.cx, xh, lane);
            }
            acc[0] = warp_sum(acc[0]);
            if (GLU) acc[1] = warp_sum(acc[1]);
            if (lane == 0) {
                float v = round_h(acc[0]);  // the GEMV output tensor is half
                if (a.bias != nullptr) v = round_h(v + h2f(bias_bits));
                if (GLU) {
                    // SiluActivation_Kernel (unary_tensor_opr.h:553) then Mul, each stored as half
                    const float g = round_h(v / (1.0f + expf(-v)));
                    v = round_h(g * round_h(acc[1]));
                }
                if (a.residual != nullptr) v = round_h(v + h2f(res_bits));
                if (TP && a.tp.push) {
                    push_y[pos][r] = __float2half_rn(v);   // sent below, R outputs at a time
                } else {
                    a.y[o] = __float2half_rn(v);
                }
            }
        }
        __syncwarp();
        if (TP && a.tp.push && !a.debug_nocompute) {
            // Tensor-parallel partial sums go straight into every rank's exchange region over NVLink. The
            // group's last warp to finish its rows sends the unit's R outputs as 8-byte stores (a 2-byte
            // store per row would cost one NVLink packet each; the words carry their own validity flag, so there is
            // nothing to publish afterwards). push_y/push_cnt are per ring slot: the slot
            // is not refilled before this warp has arrived on its empty barrier below.
            unsigned int old = 0;
            if (lane == 0) {
                __threadfence_block();
                old = atomicAdd(&push_cnt[pos], 1u);
                if (old == (unsigned)a.wpg - 1u) push_cnt[pos] = 0;
                __threadfence_block();
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == (unsigned)a.wpg - 1u) {
                // 16-byte stores of two words {half2, flag, half2, flag}: 4 outputs each, consecutive lanes write
                // consecutive addresses of one peer, so a unit costs one coalesced NVLink write per peer
                const size_t off = a.tp.dst_off + (size_t)out_base * 4;
                const int chunks = a.R >> 2;
                for (int i = lane; i < a.tp.size * chunks; i += 32) {
                    const int p = i / chunks, c = i - p * chunks;
                    const uint2 d = reinterpret_cast<const uint2 *>(push_y[pos])[c];
                    st_relaxed_sys_v4(reinterpret_cast<uint8_t *>(a.tp.peer[p]) + off + (size_t)c * 16,
                                      make_uint4(d.x, pflag, d.y, pflag));
                }
                __syncwarp();
            }
        }
        if (lane == 0) mbar_arrive(&empty_bar[pos]);
        pos += step_pos;
        phase ^= (uint32_t)step_wrap & 1u;
        if (pos >= a.nslots) { pos -= a.nslots; phase ^= 1u; }
    }
    if (tid == 0) trace_mark(a.trace, 4);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// per-device caches: a process may drive several GPUs (the engine's layer-parallel stages)
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64) ? dev : 0;
}
int sm_count() {
    static int counts[64] = {0};
    const int dev = current_device();
    if (counts[dev] == 0) {
        cudaDeviceGetAttribute(&counts[dev], cudaDevAttrMultiProcessorCount, dev);
        if (counts[dev] <= 0) counts[dev] = 148;
    }
    return counts[dev];
}

template <class Fmt, int MODE, bool GLU, bool VEC4, bool TP>
static int launch_variant_tp(const GemvArgs &a, size_t smem, int grid, bool pdl, cudaStream_t st) {
    auto kern = gemv_stream_kernel<Fmt, MODE, GLU, VEC4, TP>;
    static unsigned long long attr_done = 0ull;  // per instantiation, one bit per device: the opt-in is per device
    const int dev = current_device();
    if (!((attr_done >> dev) & 1ull)) {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done |= 1ull << dev;
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

// the tensor-parallel code (peer stores in the epilogue, the gather pre-pass) lives in its own instantiations:
// inlined into the single-GPU kernels it cost registers and made them spill
template <class Fmt, int MODE, bool GLU, bool VEC4>
static int launch_variant(const GemvArgs &a, size_t smem, int grid, bool pdl, cudaStream_t st) {
    if (a.tp.push || a.tp.gather) return launch_variant_tp<Fmt, MODE, GLU, VEC4, true>(a, smem, grid, pdl, st);
    return launch_variant_tp<Fmt, MODE, GLU, VEC4, false>(a, smem, grid, pdl, st);
}

template <bool GLU>
static int dispatch(const GemvArgs &a, int mode, bool vec4, size_t smem, int grid, bool pdl, cudaStream_t st) {
    if (mode == MODE_F16) return launch_variant<FmtQ4_B32T1, MODE_F16, GLU, false>(a, smem, grid, pdl, st);
    if (mode == MODE_HX) return launch_variant<FmtQ4_B32T1, MODE_HX, GLU, false>(a, smem, grid, pdl, st);
    switch (a.w_dtype) {
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B:
            if (vec4) return launch_variant<FmtQ4_B32T1, MODE_Q8, GLU, true>(a, smem, grid, pdl, st);
            return launch_variant<FmtQ4_B32T1, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
        case DT_Q4_B64T1: return launch_variant<FmtQ4_B64T1, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
        case DT_Q8_B32T2: return launch_variant<FmtQ8_B32T2, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
        case DT_Q6_B64T1: return launch_variant<FmtQ6_B64T1, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
        case DT_Q5_B64T1: return launch_variant<FmtQ5_B64T1, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
        case DT_Q3H_B64T1: return launch_variant<FmtQ3H_B64T1, MODE_Q8, GLU, false>(a, smem, grid, pdl, st);
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

static int *g_scratch_counter[64] = {nullptr};   // per device (a process may drive several GPUs)

int gemv_launch(const GemvLaunch &g, cudaStream_t st) {
    const int dt = g.w_dtype;
    const int cap = block_capacity(dt);
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
    a.w_gate = (const uint8_t *)g.w_gate;
    a.x = (const __half *)g.x;
    a.xq = (const uint8_t *)g.xq;
    a.norm_w = (const __half *)g.norm_w;
    a.eps = g.norm_eps;
    a.norm_base = g.norm_base;
    a.residual = (const __half *)g.residual;
    a.bias = (const __half *)g.bias;
    a.y = (__half *)g.y;
    a.trace = (unsigned long long *)g.trace;
    a.counter = (int *)g.counter;
    if (g.tp != nullptr) a.tp = *g.tp;
    a.w_table = (const uint8_t *const *)g.w_table;
    a.wg_table = (const uint8_t *const *)g.wg_table;
    a.moe_id = g.moe_id;
    IFB_CHECK_ARG(g.w_table == nullptr || (g.moe_id != nullptr && g.nseg == 1), "gemv: an expert table needs the expert id and a single segment");
    IFB_CHECK_ARG(g.w_table == nullptr || ((g.wg_table != nullptr) == (g.w_gate != nullptr)), "gemv: expert gate table and gate matrix go together");
    IFB_CHECK_ARG(!(a.tp.gather && g.xq != nullptr), "gemv: tensor-parallel gather needs an fp16 activation");
    IFB_CHECK_ARG(a.tp.size >= 1 && a.tp.size <= 8, "gemv: tensor-parallel size %d", a.tp.size);
    {
        static int nocompute = -1;
        if (nocompute < 0) { const char *e = getenv("IFB_GEMV_NOCOMPUTE"); nocompute = (e && atoi(e)) ? 1 : 0; }
        a.debug_nocompute = nocompute;
        static int late = -1;
        if (late < 0) { const char *e = getenv("IFB_PDL_LATE"); late = (e && atoi(e)) ? 1 : 0; }
        a.pdl_late = late;
        static int bc = -1;
        if (bc < 0) { const char *e = getenv("IFB_GEMV_BATCH"); bc = e ? atoi(e) : 0; }
        a.batch_claim = bc;
        static int nol2 = -1;
        if (nol2 < 0) { const char *e = getenv("IFB_NO_L2_PREFETCH"); nol2 = (e && atoi(e)) ? 1 : 0; }
        if (nol2) const_cast<GemvLaunch &>(g).l2_prefetch = 0;
    }

    // rows per unit: R*row_bytes must be a multiple of 16 (bulk-copy granularity) and R must
    // divide every segment; prefer 8 (one row per consumer warp of a group), 4 when an 8-row slot
    // would be too fat, 16/32 for very short rows.
    const int nmat = g.w_gate ? 2 : 1;
    int gcd_rows = 0;
    for (int i = 0; i < g.nseg; i++) {
        IFB_CHECK_ARG(g.rows[i] > 0 && g.w[i] != nullptr, "gemv: empty segment %d", i);
        IFB_CHECK_ARG(((uintptr_t)g.w[i] & 15) == 0, "gemv: weight pointer not 16-byte aligned");
        int x = gcd_rows, y = g.rows[i];
        while (y) { int t = x % y; x = y; y = t; }
        gcd_rows = gcd_rows == 0 ? g.rows[i] : x;
    }
    IFB_CHECK_ARG(g.w_gate == nullptr || ((uintptr_t)g.w_gate & 15) == 0, "gemv: gate weight pointer not 16-byte aligned");
    auto valid = [&](int c) { return gcd_rows % c == 0 && ((int64_t)c * a.row_bytes) % 16 == 0; };
    auto sbytes = [&](int c) { return (int64_t)c * a.row_bytes * nmat; };
    int R = 0;
    if (valid(8) && sbytes(8) <= 32 * 1024) R = 8;
    else if (valid(4)) R = 4;
    else if (valid(8)) R = 8;
    if (R == 8 && valid(16) && sbytes(16) <= 16 * 1024) R = 16;   // very short rows: fatter copies
    if (R == 16 && valid(32) && sbytes(32) <= 16 * 1024) R = 32;
    {
        static int env_r = -1;        // tuning aid: IFB_GEMV_R forces the rows per unit when it tiles the launch
        if (env_r < 0) { const char *e = getenv("IFB_GEMV_R"); env_r = e ? atoi(e) : 0; }
        if (env_r > 0 && valid(env_r)) R = env_r;
    }
    IFB_CHECK_ARG(R != 0, "gemv: rows (gcd %d) x row_bytes %d cannot be tiled in 16-byte units "
                  "(rows must be a multiple of 4, or 8 for 34-byte blocks)", gcd_rows, a.row_bytes);
    a.R = R;
    a.wpg = R < GEMV_CWARPS ? R : GEMV_CWARPS;

    int units = g.rows[0] / R;
    a.w0 = (const uint8_t *)g.w[0];
    a.bytes0 = (long long)g.rows[0] * a.row_bytes;
    a.w1 = a.w2 = nullptr;
    a.bytes1 = a.bytes2 = 0;
    a.unit1 = a.unit2 = 0x7fffffff;
    if (g.nseg > 1) { a.w1 = (const uint8_t *)g.w[1]; a.bytes1 = (long long)g.rows[1] * a.row_bytes; a.unit1 = units; units += g.rows[1] / R; }
    if (g.nseg > 2) { a.w2 = (const uint8_t *)g.w[2]; a.bytes2 = (long long)g.rows[2] * a.row_bytes; a.unit2 = units; units += g.rows[2] / R; }
    a.units = units;
    {   // software prefetch distance (bytes -> units); 0 disables
        static long long dist_bytes = -1;
        if (dist_bytes < 0) {
            const char *e = getenv("IFB_L2_DIST_MB");
            dist_bytes = (e ? atoll(e) : 0) * 1024 * 1024;   // measured on B200: no gain (profiles/README.md), off by default
        }
        const long long unit_bytes = (long long)R * a.row_bytes * nmat;
        a.l2_prefetch = g.l2_prefetch ? (int)(dist_bytes / unit_bytes) : 0;
    }

    const bool vec4 = mode == MODE_Q8 && (dt == DT_Q4_B32T1A || dt == DT_Q4_B32T1B) && g.cx % 128 == 0;

    const int64_t slot_bytes = (int64_t)R * a.row_bytes * nmat;
    a.slot_stride = (int)((slot_bytes + 16 + 127) / 128 * 128);
    a.nxb_alloc = (int)(g.cx / 32);
    if (vec4) a.nxb_alloc = (a.nxb_alloc + 127) / 128 * 128;   // x_slot<true>() scatters inside passes of 128 blocks
    int64_t xbytes = (mode == MODE_Q8) ? (int64_t)a.nxb_alloc * 40 : g.cx * 2;
    xbytes = (xbytes + 127) / 128 * 128;
    a.xg_off = (int)xbytes;
    if (a.tp.gather) xbytes += g.cx * 2;      // the gathered fp16 vector, consumed by the staging pass
    const int64_t xpad = (xbytes + 127) / 128 * 128;
    // default: at most ~half of the SM's shared memory, so two kernels of a PDL chain co-reside
    static int64_t env_budget = -1;   // tuning aid: IFB_GEMV_SMEM_KB overrides the default per-CTA ring budget
    if (env_budget < 0) { const char *e = getenv("IFB_GEMV_SMEM_KB"); env_budget = e ? atoll(e) * 1024 : 0; }
    // Default ring: (almost) the whole SM. Measured on B200 (tools/sweep_smem.sh, Llama-2-7B Q4 decode): 110 KB
    // (two kernels of a PDL chain co-resident) 512 tok/s, 150 KB 559, 200-216 KB 586 -- bytes in flight per SM
    // matter more than the next kernel's early ring fill. Callers that want a co-resident successor (the QKV
    // GEMV ahead of attention) pass a smaller budget.
    int64_t budget = (g.smem_budget > 0 ? g.smem_budget : (env_budget > 0 ? env_budget : 216 * 1024));
    int groups = GEMV_CWARPS / a.wpg;
    int nslots = (int)((budget - xpad) / a.slot_stride);
    if (nslots < 2) {
        budget = 216 * 1024;
        nslots = (int)((budget - xpad) / a.slot_stride);
    }
    IFB_CHECK_ARG(nslots >= 1, "gemv: a %lld-byte slot does not fit shared memory", (long long)slot_bytes);
    if (nslots > GEMV_MAX_SLOTS) nslots = GEMV_MAX_SLOTS;
    // A consumer group must meet EVERY fill of the slots it reads (mbarrier parity cannot tell fill k
    // from fill k+2): slots are owned by groups, nslots is a multiple of the group count.
    if (groups > nslots) groups = nslots;
    nslots = nslots / groups * groups;
    a.groups = groups;
    a.nslots = nslots;
    const size_t smem = (size_t)nslots * a.slot_stride + xpad;

    int grid = sm_count();
    if (grid > units) grid = units;
    if (a.counter == nullptr) {   // stand-alone call: a library-owned cursor, zeroed on the caller's stream
        int *&sc = g_scratch_counter[current_device()];
        if (sc == nullptr) IFB_CUDA(cudaMalloc((void **)&sc, 256));
        IFB_CUDA(cudaMemsetAsync(sc, 0, sizeof(int), st));
        a.counter = sc;
    }

    if (g.w_gate) return dispatch<true>(a, mode, vec4, smem, grid, g.pdl, st);
    return dispatch<false>(a, mode, vec4, smem, grid, g.pdl, st);
}

}  // namespace ifb
