// One decoded token = ONE persistent kernel launch, sm_100a.
//
// Replaces, for batch-1 decode, the whole of GpuInferenceWorker::Run -> ProcessGpuLayer ->
// {ProcessGpuLayer_Attention, ProcessGpuLayer_FeedForward} -> ProcessPostLayer (reference
// src/transformer/inference_worker.cc:234-977, 983-1724, 1726-1900, 552-640) and, under tensor parallelism,
// DistributeAndMergeTensors (:2148-2335). The reference issues ~60 synchronous launches per layer; round 1 of this
// runtime issued 5 fused kernels per layer and measured that what was left was FIXED COST PER LAUNCH (activation
// staging behind a cold ring, ring refill after the grid dependency, tail): 4-5 us on launches whose weight stream
// takes 1.6-8.6 us. Here the launch boundaries are gone:
//
//  * grid = one CTA per SM, resident for the whole token: 15 consumer warps + 1 producer warp;
//  * the PRODUCER (one elected thread; shared-ring schedule, PRIV = false) walks every op of the token in order --
//    per layer wq|wk|wv, the KV cache, wo, w1|w3, w2, then the lm_head -- claiming units from per-op global cursors
//    and streaming them with 1-D TMA bulk copies into a shared-memory ring. It never waits for an activation: while
//    the consumers of all SMs synchronise at an op boundary, the HBM pipe keeps filling the rings with the NEXT op's
//    weights (the rings of 148 SMs hold ~27 MB = 4 us of HBM time);
//  * the CONSUMERS take ring slots in order; a slot's tasks (2 adjacent rows / two w1,w3 row pairs / 32 cache
//    positions) are dealt round-robin to the 15 warps. At an op boundary every CTA stages the new op's input into
//    shared memory itself: gather of what the previous op left -- as FLAG-IN-DATA WORDS {half2, flag} that the reader
//    polls (q|k|v, attention output, h, the wo / w2 outputs; no fence, no atomic), or behind a completion count where
//    a reduction needs one (attention tiles -> merge) -- then residual add (fp16, the reference's order), RMSNorm,
//    Q8_B32T2 quantization (the reference's alg-2 arithmetic), or RoPE + two-term int8 q for attention;
//  * PRIV = true (IFB_STEP_PRIV=1): no producer warp -- every consumer warp streams its own statically assigned tasks
//    through two private slots (see the kernel body);
//  * MEASURED (profiles/r2_ncu.md): a boundary costs ~4 us, not the ~1 us this design assumed (tail of the
//    producing op + store -> poll round trip under streaming load + staging with two CTA barriers + first task): 5 per
//    layer are 22 us beside 34 us of streaming on the 7B model; neither schedule is instruction-bound;
//  * attention is flash-decoding over the cache as stored: a task is one 32-position tile of one kv head, its
//    (m, l, acc) partial goes to an L2-resident workspace, one CTA per head merges them (log-sum-exp) before wo;
//  * tensor parallelism: the row-parallel GEMVs (wo, w2) store their partial sums straight into every peer's
//    exchange region over NVLink as flag-in-data words {half2, flag}; the next op's staging polls exactly those
//    words (bounded: a deadline, then an error word in pinned host memory and a trap -- never an endless spin)
//    and adds them in rank order with fp16 rounding (MergeTensors, inference_worker.cc:2264-2289).
//
// The arithmetic of every GEMV row and of the staging is the code of ifb_gemv_dev.cuh, shared with the stand-alone
// kernels: the step kernel and the operator entries produce the same bits.
#include <cstdlib>

#include "ifb_gemv_dev.cuh"
#include "ifb_kernels.h"
#include "ifb_step.h"

namespace ifb {

// 15 consumer warps + 1 producer warp = 16 warps, four per SM sub-partition: 128 registers per thread are available.
// (16 + 1 warps put five warps on one sub-partition: 96 registers, and the compiler kept the ring-walk state of the
// consumer loop in local memory -- reloaded for every slot, with next to no L1 behind it.)
constexpr int STEP_CWARPS = 15;
constexpr int STEP_CTHREADS = STEP_CWARPS * 32;
constexpr int STEP_THREADS = STEP_CTHREADS + 32;
constexpr int STEP_MAX_SLOTS = 16;
constexpr int STEP_MAX_BARS = 2 * STEP_CWARPS + 2;   // warp-private pipelines: two slots (and barriers) per consumer warp
constexpr int STEP_PRIV_CHUNK = 5120;                // payload bytes of a private slot: two Q4_B32T1 rows of 4096 columns
constexpr int STEP_TARGET_UNIT = 20480;   // bytes per ring slot the unit sizes aim for

struct FmtF16 {   // tag: fp16 weights (tensors below tensor_quant_threshold, network_builder.cc:1558)
    static constexpr bool XSPLIT = false;
};
template <class Fmt> struct is_f16_fmt { static constexpr bool value = false; };
template <> struct is_f16_fmt<FmtF16> { static constexpr bool value = true; };
template <class Fmt> struct is_q4b32_fmt { static constexpr bool value = false; };
template <> struct is_q4b32_fmt<FmtQ4_B32T1> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_cg128(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cg_f32(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg_f32x4(const float *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v2(void *p, uint2 v) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void step_abort(const StepParams &P, int code, int op, int detail) {
    if (P.err != nullptr) {
        P.err[1] = op; P.err[2] = detail; P.err[3] = (int)blockIdx.x;
        P.err[0] = code;
        __threadfence_system();
    }
    __trap();
}

// every wait of this kernel is bounded: past the deadline the kernel leaves an error record and traps. The threads
// sleep in the barrier unit (suspend-time hint) rather than spin: a spinning consumer warp costs the working warps issue
// slots (ncu: the first version spent 5 % of all executed instructions in this loop).
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar_addr, uint32_t parity, const StepParams &P, int op, unsigned pause_ns = 100u) {
    if (mbar_try_wait_a(bar_addr, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned it = 0;
    while (!mbar_try_wait_a(bar_addr, parity)) {
        __nanosleep(pause_ns);     // a waiting warp must not eat the working warps' issue slots (ncu: 8 % of all instructions)
        if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_RING, op, (int)parity);
    }
}

// grid-wide dependency: thread 0 of the consumers waits until `target` CTAs have released the counter, then the
// consumer warps meet at their named barrier. Data published before the release is read with ld.global.cg.
__device__ __forceinline__ void dep_wait(const unsigned int *ctr, unsigned int target, const StepParams &P, int op, int tid) {
    if (tid == 0) {
        if (ld_acquire_gpu_u32(ctr) < target) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned it = 0;
            while (ld_acquire_gpu_u32(ctr) < target) {
                if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_DEP, op, (int)ld_acquire_gpu_u32(ctr));
            }
        }
    }
    named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
}
__device__ __forceinline__ void dep_signal(unsigned int *ctr, int tid) {
    named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);     // every consumer thread's stores are ordered before thread 0's fence
    if (tid == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
    }
}

__device__ __forceinline__ void trace_op(const StepParams &P, int op, int k, int tid) {
    if (P.trace != nullptr && tid == 0) P.trace[((size_t)op * gridDim.x + blockIdx.x) * 8 + k] = globaltimer_ns();
}

// monotone key of (logit, index): larger logit wins, ties go to the smaller index (argmax: first max wins)
// profile counters of the timeline probe (IFB_TRACE=1): entries 4..7 of an (op, CTA) record
__device__ __forceinline__ void trace_add(const StepParams &P, int op, int k, unsigned long long v) {
    if (P.trace != nullptr) P.trace[((size_t)op * gridDim.x + blockIdx.x) * 8 + k] += v;
}
__device__ __forceinline__ long long clk() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

__device__ __forceinline__ unsigned long long amax_key(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

// ---------------------------------------------------------------------------------------
// flag-in-data vectors: every 8-byte word is {two fp16 values, 32-bit flag}; a word is valid for (step, tag) when its
// flag equals (step sequence number << 8 | tag). 8-byte stores are atomic: the writer needs no fence, the reader polls
// the words it needs -- the data IS the signal (one L2 round trip where a completion count costs three: the writer's
// release add, the reader's poll, then the load).
// ---------------------------------------------------------------------------------------
// 8 consecutive fp16 values = words 4g..4g+3 of `words`, polled until valid
// (poll = false: a partial issue -- measurement aid -- in which nobody writes the words: take what is there)
__device__ __forceinline__ uint4 ll_load8(const uint4 *words, int g, uint32_t flag, const StepParams &P, int op, bool poll = true) {
    const uint4 *src = words + 2 * g;
    uint4 w0 = ld_relaxed_sys_v4(src), w1 = ld_relaxed_sys_v4(src + 1);
    if (poll && (w0.y != flag || w0.w != flag || w1.y != flag || w1.w != flag)) {
        const unsigned long long t0 = globaltimer_ns();
        unsigned it = 0;
        while (w0.y != flag || w0.w != flag || w1.y != flag || w1.w != flag) {
            w0 = ld_relaxed_sys_v4(src);
            w1 = ld_relaxed_sys_v4(src + 1);
            if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_PEER, op, g);
        }
    }
    return make_uint4(w0.x, w0.z, w1.x, w1.z);
}

// Staging of a GEMV input from values that arrive in REGISTERS (load(g) -> the 8 fp16 values of group g, already final):
// [RMSNorm with the sum of squares taken from the registers] -> Q8_B32T2 quantization -> the staged layout. Two
// barriers (sum of squares, done) and no shared-memory round trip of the vector. Bit-for-bit the arithmetic of
// stage_x_q8 (ifb_gemv_dev.cuh). A thread owns groups tid, tid + 480, tid + 960 of a round; vectors with a norm fit
// one round (cx <= 11520, checked when the plan is built).
template <bool XSPLIT, class LoadF>
__device__ __forceinline__ void stage_fused_q8(const StepOpShape &s, const __half *norm_w, float eps, uint8_t *xst, float *red,
                                               int tid, LoadF load) {
    constexpr int NT = STEP_CTHREADS, KPT = 3;
    const int warp = tid >> 5, lane = tid & 31;
    const int ng = s.cx >> 3;
    uint8_t *xlo = xst, *xhi = xst + (size_t)s.nxb * 16;
    float2 *xs = reinterpret_cast<float2 *>(xst + (size_t)s.nxb * 32);
    const bool norm = norm_w != nullptr, vec4 = s.vec4 != 0;
    const uint4 *w4 = reinterpret_cast<const uint4 *>(norm_w);
    float rs = 1.f;
    for (int g0 = 0; g0 < ng; g0 += NT * KPT) {
        const int ga = g0 + tid, gb = ga + NT, gc = gb + NT;
        uint4 wv0 = make_uint4(0, 0, 0, 0), wv1 = wv0, wv2 = wv0, xv0 = wv0, xv1 = wv0, xv2 = wv0;
        if (norm) {   // the norm weights do not depend on the previous op: their round trip runs under the wait for x
            if (ga < ng) wv0 = w4[ga];
            if (gb < ng) wv1 = w4[gb];
            if (gc < ng) wv2 = w4[gc];
        }
        if (ga < ng) xv0 = load(ga);
        if (gb < ng) xv1 = load(gb);
        if (gc < ng) xv2 = load(gc);
        if (norm) {
            // Tensor_RmsNorm_Kernel (unary_tensor_opr.h:217): scale = 1/sqrt(mean(x^2) + eps)
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
            for (int i = 0; i < STEP_CWARPS; i++) t += red[i];
            rs = 1.0f / sqrtf(t / (float)s.cx + eps);
        }
#pragma unroll 1
        for (int k = 0; k < KPT; k++) {
            const int g = g0 + k * NT + tid;   // a quad of lanes shares one block
            if (g0 + k * NT + warp * 32 >= ng) break;   // warp-uniform: nothing left for this warp
            const uint4 xk = k == 0 ? xv0 : (k == 1 ? xv1 : xv2);
            const uint4 wk = k == 0 ? wv0 : (k == 1 ? wv1 : wv2);
            float v[8];
            {
                const __half2 *h = reinterpret_cast<const __half2 *>(&xk);
                const __half2 *hw = reinterpret_cast<const __half2 *>(&wk);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 f = __half22float2(h[i]);
                    if (norm) {  // v = x*scale*(multi_base + w), stored as half by the reference's RmsNorm (multi_base 0)
                        const float2 fw = __half22float2(hw[i]);
                        f.x = round_h(f.x * rs * (0.f + fw.x));
                        f.y = round_h(f.y * rs * (0.f + fw.y));
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
                const int b = g >> 2, j = g & 3;
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

// ---------------------------------------------------------------------------------------
// residual stream: resid (shared memory, fp16 [d]) += the row-parallel GEMV's output
// ---------------------------------------------------------------------------------------
// single GPU: y = round_h(W x) was stored by the consumers of all CTAs (half tensor, as the reference's GEMV
// output), x_new = round_h(y + resid) (TensorOpr::Add, inference_worker.cc:849)
__device__ __forceinline__ void resid_add_local(const StepParams &P, uint4 *resid, int tid) {
    const int ng = P.d >> 3;
    for (int g = tid; g < ng; g += STEP_CTHREADS) {
        const uint4 yv = ld_cg128(reinterpret_cast<const uint4 *>(P.ybuf) + g);
        const uint4 rv = resid[g];
        uint4 out;
        const __half2 *y = reinterpret_cast<const __half2 *>(&yv), *r = reinterpret_cast<const __half2 *>(&rv);
        __half2 *o = reinterpret_cast<__half2 *>(&out);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 yf = __half22float2(y[i]), rf = __half22float2(r[i]);
            o[i] = __floats2half2_rn(yf.x + rf.x, yf.y + rf.y);
        }
        resid[g] = out;
    }
}

// tensor parallel: the partial sums of all ranks arrive in this rank's exchange region as flag-in-data words
// {half2, flag}; they are added in rank order with fp16 rounding after every add, the residual last
// (MergeTensors adds worker 1..n-1 onto worker 0 with TensorOpr::Add, inference_worker.cc:2264-2289, then :849)
__device__ __forceinline__ void resid_add_gather(const StepParams &P, uint4 *resid, uint32_t idx, uint32_t seq, int op, int tid) {
    const uint32_t flag = (seq << 8) | idx;
    const uint8_t *base = reinterpret_cast<const uint8_t *>(P.peer[P.tp_rank]) + P.x_recv + (long long)(idx & 1u) * P.x_recv_parity;
    const int ng = P.d >> 3;
    for (int g = tid; g < ng; g += STEP_CTHREADS) {
        float acc[8];
        for (int p = 0; p < P.tp_size; p++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(base + (size_t)p * P.x_recv_rank) + 2 * g;
            uint4 w0 = ld_relaxed_sys_v4(src), w1 = ld_relaxed_sys_v4(src + 1);
            if (w0.y != flag || w0.w != flag || w1.y != flag || w1.w != flag) {
                const unsigned long long t0 = globaltimer_ns();
                unsigned it = 0;
                while (w0.y != flag || w0.w != flag || w1.y != flag || w1.w != flag) {
                    w0 = ld_relaxed_sys_v4(src);
                    w1 = ld_relaxed_sys_v4(src + 1);
                    if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_PEER, op, p);
                }
            }
            const uint32_t dw[4] = {w0.x, w0.z, w1.x, w1.z};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&dw[k]));
                if (p == 0) { acc[2 * k] = f.x; acc[2 * k + 1] = f.y; }
                else { acc[2 * k] = round_h(acc[2 * k] + f.x); acc[2 * k + 1] = round_h(acc[2 * k + 1] + f.y); }
            }
        }
        const uint4 rv = resid[g];
        uint4 out;
        __half2 *o = reinterpret_cast<__half2 *>(&out);
        const __half2 *r = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 rf = __half22float2(r[i]);
            o[i] = __floats2half2_rn(acc[2 * i] + rf.x, acc[2 * i + 1] + rf.y);
        }
        resid[g] = out;
    }
}

// ---------------------------------------------------------------------------------------
// attention: staging of q and the new K/V row, the 32-position task, the merge
// ---------------------------------------------------------------------------------------
// staged layout inside the x area (HD = head_dim, NB = HD/32):
//   Q8 cache : int8 qh[heads][HD] | int8 ql[heads][HD] | float qs[heads][NB][2]
//   F16 cache: float q[heads][HD/16][20]
struct AttLayout {
    int qh, ql, qs, total;
};
__host__ __device__ inline AttLayout att_layout(int heads, int hd, bool q8) {
    AttLayout a;
    const int nb = hd / 32;
    if (q8) {
        a.qh = 0; a.ql = heads * hd; a.qs = 2 * heads * hd;
        a.total = a.qs + heads * nb * 8;
    } else {
        a.qh = 0; a.ql = 0; a.qs = 0;
        a.total = heads * (hd / 16) * 20 * 4;
    }
    a.total = (a.total + 127) / 128 * 128;
    return a;
}

// RoPE of 8 consecutive dims d0..d0+7 of one head row (values v, partner group o for the half-split order), each
// result rounded to half as PositionEmbedding stores it (tensor_opr.cu:693-760, unary_tensor_opr.h:662-740)
__device__ __forceinline__ void rope8(float *v, const float *o, int d0, int hd, int order, const float *s_cos, const float *s_sin) {
    if (order == 2) {
        const bool first = d0 < hd / 2;
        const int p0 = first ? d0 : d0 - hd / 2;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float cs = s_cos[p0 + i], sn = s_sin[p0 + i];
            v[i] = first ? round_h(v[i] * cs - o[i] * sn) : round_h(o[i] * sn + v[i] * cs);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float cs = s_cos[(d0 + i) >> 1], sn = s_sin[(d0 + i) >> 1];
            const float a = v[i], b = v[i + 1];
            v[i] = round_h(a * cs - b * sn);
            v[i + 1] = round_h(a * sn + b * cs);
        }
    }
}
__device__ __forceinline__ void unpack8(const uint4 w, float *v) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&w);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ float quad_max(float m) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    return fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
}
__device__ __forceinline__ uint2 pack8_s8(const int *q) {
    return make_uint2((q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24),
                      (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24));
}

// q of every head of this rank: RoPE, round to half (PositionEmbedding stores half), scale; for a Q8 cache the two
// int8 terms per 32-dim block, q ~= s1*qh + s2*ql (s2 <= s1/254). One thread = 8 consecutive dims, four neighbouring
// lanes = one block (its max by two shuffles).
template <bool Q8>
__device__ __forceinline__ void stage_att(const StepParams &P, uint8_t *xst, const float *s_cos, const float *s_sin, int tid,
                                          uint32_t flag, int op) {
    const int hd = P.hd, nb = hd / 32;
    const AttLayout L = att_layout(P.heads, hd, Q8);
    const uint4 *qw = reinterpret_cast<const uint4 *>(P.qkvw);     // the wq|wk|wv GEMV wrote {half2, flag} words
    const int nq = P.q_dim >> 3;
    const bool half_split = P.rope_order == 2;
    for (int base = 0; base < nq; base += STEP_CTHREADS) {
        if (base + (tid & ~31) >= nq) break;          // nothing left for this warp
        const int g8 = base + tid;
        const bool aq = g8 < nq;
        const int e0 = aq ? g8 * 8 : 0;
        const int hq = e0 / hd, dq = e0 - hq * hd;
        uint4 wq_ = make_uint4(0, 0, 0, 0), wqo = wq_;
        if (aq) {
            const bool poll = (P.parts & STEP_PART_GEMV) != 0;
            wq_ = ll_load8(qw, g8, flag, P, op, poll);
            if (half_split) wqo = ll_load8(qw, (hq * hd + (dq < hd / 2 ? dq + hd / 2 : dq - hd / 2)) >> 3, flag, P, op, poll);
        }
        float x[8], o[8];
        unpack8(wq_, x);
        unpack8(wqo, o);
        rope8(x, o, dq, hd, P.rope_order, s_cos, s_sin);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] *= P.softmax_scale;
        if (Q8) {
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) m = fmaxf(m, fabsf(x[i]));
            // (the two-term split is this kernel's own representation of q, not a reference format: any rounding of
            // x/s1 within the int8 range serves -- the second term carries the remainder -- so reciprocal multiplies)
            const float s1 = quad_max(m) * (1.0f / 127.0f);
            const float i1 = s1 > 0.f ? __frcp_rn(s1) : 0.f;
            int qh[8], ql[8];
            float r[8], m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float t = fminf(fmaxf(rintf(x[i] * i1), -127.f), 127.f);
                qh[i] = (int)t;
                r[i] = x[i] - t * s1;
                m2 = fmaxf(m2, fabsf(r[i]));
            }
            const float s2 = quad_max(m2) * (1.0f / 127.0f);
            const float i2 = s2 > 0.f ? __frcp_rn(s2) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) ql[i] = (int)fminf(fmaxf(rintf(r[i] * i2), -127.f), 127.f);
            if (aq) {
                *reinterpret_cast<uint2 *>(xst + L.qh + e0) = pack8_s8(qh);
                *reinterpret_cast<uint2 *>(xst + L.ql + e0) = pack8_s8(ql);
                if ((dq & 31) == 0) reinterpret_cast<float2 *>(xst + L.qs)[hq * nb + (dq >> 5)] = make_float2(s1, s2);
            }
        } else if (aq) {
            float *dst = reinterpret_cast<float *>(xst) + ((size_t)hq * (hd / 16) + (dq >> 4)) * 20 + (dq & 15);
            *reinterpret_cast<float4 *>(dst) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
        }
    }
}

// signed byte k of w -> float without I2F (see ifb_attention.cu)
__device__ __forceinline__ float s8f_(uint32_t w_biased, int k) {
    const uint32_t bits = __byte_perm(w_biased, 0x4b000000u, 0x7650u + (uint32_t)k);
    return __uint_as_float(bits) - 8388736.0f;
}

// One task = the 32 positions [pbase, pbase+32) of kv head kvh against the q heads of its group, one q head after the
// other (the tile is read from shared memory once per q head: grouped-query models have little cache to stream, the
// op stays HBM-bound); the warp writes the tile's (m, l, acc[HD]) partial of each q head. The arithmetic is the scan
// body of attn_decode_kernel (ifb_attention.cu): dp4a scores with the two-term int8 q for a Q8 cache, fp32 softmax,
// lane-owns-4-dims P.V.
template <int HD, bool Q8>
__device__ __forceinline__ void att_task(const StepParams &P, const StepLayer &lw, const float *s_cos, const float *s_sin, uint32_t qflag,
                                         uint8_t *kslot, uint8_t *vslot, int UP, int tt, int kvh, int p0, int ctx, const uint8_t *xst, float *sp, int lane) {
    constexpr int NB = HD / 32;
    const int pbase = p0 + tt * 32;
    if (pbase >= ctx) return;
    const AttLayout L = att_layout(P.heads, HD, Q8);
    const int G = P.group;
    // kslot / vslot: UP positions of the head's K / V plane as the bulk copies left them (Q8: codes, then block scales)
    uint8_t *tK, *tKs = nullptr, *tV, *tVs = nullptr;
    if (Q8) {
        tK = kslot + tt * 32 * HD;
        tKs = kslot + UP * HD + tt * 32 * NB * 2;
        tV = vslot + tt * 32 * HD;
        tVs = vslot + UP * HD + tt * 32 * NB * 2;
    } else {
        tK = kslot;
        tV = vslot;
    }
    if (ctx - 1 >= pbase && ctx - 1 < pbase + 32) {
        // This tile holds the position being decoded. Its bulk copy was issued before this token's K/V row existed
        // (or raced with it): the warp builds the row here -- RoPE on k, then the cache's storage form (SetKRows /
        // SetVRows, kv_cache.cc:104-183; Q8: Tensor_QuantizeQ8_B32T2_Alg2_Kernel, tensor_quant.h:45-82) -- stores
        // it to the cache and puts it in place in the staged tile. One task per kv head and token does this.
        constexpr int DPL = HD / 32;                       // dims per lane
        const int r = (ctx - 1) - pbase, d0 = lane * DPL;
        float kx[DPL], vx[DPL];
        {
            // this token's k and v of the head, as the wq|wk|wv GEMV left them: {half2, flag} words, polled
            const uint2 *kwp = P.qkvw + ((P.q_dim + kvh * HD + d0) >> 1), *vwp = kwp + (P.kv_dim >> 1);
            uint32_t kw[DPL / 2], vw[DPL / 2];
#pragma unroll
            for (int i = 0; i < DPL / 2; i++) {
                uint2 a_ = ld_relaxed_sys_v2(kwp + i), b_ = ld_relaxed_sys_v2(vwp + i);
                if ((P.parts & STEP_PART_GEMV) && (a_.y != qflag || b_.y != qflag)) {
                    const unsigned long long t0 = globaltimer_ns();
                    unsigned it = 0;
                    while (a_.y != qflag || b_.y != qflag) {
                        a_ = ld_relaxed_sys_v2(kwp + i);
                        b_ = ld_relaxed_sys_v2(vwp + i);
                        if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_PEER, -1, kvh);
                    }
                }
                kw[i] = a_.x; vw[i] = b_.x;
            }
#pragma unroll
            for (int i = 0; i < DPL / 2; i++) {
                const float2 fk = __half22float2(*reinterpret_cast<const __half2 *>(&kw[i]));
                const float2 fv = __half22float2(*reinterpret_cast<const __half2 *>(&vw[i]));
                kx[2 * i] = fk.x; kx[2 * i + 1] = fk.y;
                vx[2 * i] = fv.x; vx[2 * i + 1] = fv.y;
            }
        }
        if (P.rope_order == 2) {
            // half-split pairs (d, d + HD/2): the partner sits 16 lanes away
            const bool first = d0 < HD / 2;
#pragma unroll
            for (int i = 0; i < DPL; i++) {
                const float o = __shfl_xor_sync(0xffffffffu, kx[i], 16);
                const int pi = (first ? d0 : d0 - HD / 2) + i;
                const float cs = s_cos[pi], sn = s_sin[pi];
                kx[i] = first ? round_h(kx[i] * cs - o * sn) : round_h(o * sn + kx[i] * cs);
            }
        } else {
#pragma unroll
            for (int i = 0; i < DPL; i += 2) {
                const float cs = s_cos[(d0 + i) >> 1], sn = s_sin[(d0 + i) >> 1];
                const float a_ = kx[i], b_ = kx[i + 1];
                kx[i] = round_h(a_ * cs - b_ * sn);
                kx[i + 1] = round_h(a_ * sn + b_ * cs);
            }
        }
        const int64_t row = (int64_t)kvh * P.max_ctx_pad + (ctx - 1);
#pragma unroll
        for (int which = 0; which < 2; which++) {
            float *x = which == 0 ? kx : vx;
            uint8_t *cache = which == 0 ? lw.kcache : lw.vcache;
            uint8_t *trow = (which == 0 ? tK : tV) + r * (Q8 ? HD : HD * 2);
            if (Q8) {
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < DPL; i++) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
                for (int of = 1; of < 32 / DPL; of <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));   // the block's 32/DPL lanes
                const float scale = __fdiv_rn(m, 127.0f);
                const bool dead = scale <= 0.000001f;
                const float inv = dead ? 0.f : __frcp_rn(scale);
                uint32_t pk = 0;
#pragma unroll
                for (int i = 0; i < DPL; i++) {
                    const int t = dead ? 0 : q8_round_div(x[i], scale, inv);   // == (int)roundf(x / scale), see ifb_common.cuh
                    pk |= (uint32_t)(min(max(t, -128), 127) & 0xff) << (8 * i);
                }
                const __half hs = __float2half_rn(scale);
                if (DPL == 4) {
                    reinterpret_cast<uint32_t *>(trow)[lane] = pk;
                    reinterpret_cast<uint32_t *>(cache + row * HD)[lane] = pk;
                } else {
                    reinterpret_cast<uint16_t *>(trow)[lane] = (uint16_t)pk;
                    reinterpret_cast<uint16_t *>(cache + row * HD)[lane] = (uint16_t)pk;
                }
                if ((d0 & 31) == 0) {
                    reinterpret_cast<__half *>(which == 0 ? tKs : tVs)[r * NB + (d0 >> 5)] = hs;
                    reinterpret_cast<__half *>(cache + (int64_t)P.kv_heads * P.max_ctx_pad * HD)[row * NB + (d0 >> 5)] = hs;
                }
            } else {
#pragma unroll
                for (int i = 0; i < DPL; i += 2) {
                    const __half2 hv = __floats2half2_rn(x[i], x[i + 1]);
                    const uint32_t bits = *reinterpret_cast<const uint32_t *>(&hv);
                    reinterpret_cast<uint32_t *>(trow)[(d0 + i) >> 1] = bits;
                    reinterpret_cast<uint32_t *>(cache + row * HD * 2)[(d0 + i) >> 1] = bits;
                }
            }
        }
        fence_proxy_async_smem();   // generic-proxy writes into a buffer the next bulk copy (async proxy) will overwrite
        __syncwarp();
    }
    const int tile = pbase >> 5;
    const int npos = min(32, ctx - pbase);
    for (int g = 0; g < G; g++) {
        const int hq = kvh * G + g;
        // ---- scores
        if (Q8) {
            constexpr int C = HD / 16;
            constexpr int RPL = (C >= 8) ? 1 : 8 / C;
            const int rot = lane / RPL;
            float sc = 0.f;
            const float2 *qs = reinterpret_cast<const float2 *>(xst + L.qs) + hq * NB;
            const uint8_t *qh_ = xst + L.qh + (size_t)hq * HD, *ql_ = xst + L.ql + (size_t)hq * HD;
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int ci = (j + rot) & (C - 1);
                const uint4 kw = *reinterpret_cast<const uint4 *>(tK + lane * HD + ci * 16);
                const float ksc = __half2float(reinterpret_cast<const __half *>(tKs)[lane * NB + (ci >> 1)]);
                const uint4 qh = *reinterpret_cast<const uint4 *>(qh_ + ci * 16);
                const uint4 ql = *reinterpret_cast<const uint4 *>(ql_ + ci * 16);
                int dh = dp4a_ss(kw.x, qh.x, 0), dl = dp4a_ss(kw.x, ql.x, 0);
                dh = dp4a_ss(kw.y, qh.y, dh); dl = dp4a_ss(kw.y, ql.y, dl);
                dh = dp4a_ss(kw.z, qh.z, dh); dl = dp4a_ss(kw.z, ql.z, dl);
                dh = dp4a_ss(kw.w, qh.w, dh); dl = dp4a_ss(kw.w, ql.w, dl);
                const float2 s12 = qs[ci >> 1];
                const float t = fmaf(s12.x, (float)dh, s12.y * (float)dl);
                sc = fmaf(ksc, t, sc);
            }
            sp[lane] = (pbase + lane < ctx) ? sc : -INFINITY;
        } else {
            constexpr int LPP = HD / 16, PPW = 32 / LPP;
            const int sub = lane % LPP, pl = lane / LPP;
            const float4 *qv = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(xst) + ((size_t)hq * (HD / 16) + sub) * 20);
            const float4 q0 = qv[0], q1 = qv[1], q2 = qv[2], q3 = qv[3];
#pragma unroll 2
            for (int pp = pl; pp < 32; pp += PPW) {
                float kf[16];
                unpack8(*reinterpret_cast<const uint4 *>(tK + (pp * HD + sub * 16) * 2), kf);
                unpack8(*reinterpret_cast<const uint4 *>(tK + (pp * HD + sub * 16) * 2 + 16), kf + 8);
                float s0 = 0.f, s1 = 0.f;   // two chains, in the order of attn_decode_kernel
                s0 = fmaf(kf[0], q0.x, s0); s1 = fmaf(kf[4], q1.x, s1);
                s0 = fmaf(kf[1], q0.y, s0); s1 = fmaf(kf[5], q1.y, s1);
                s0 = fmaf(kf[2], q0.z, s0); s1 = fmaf(kf[6], q1.z, s1);
                s0 = fmaf(kf[3], q0.w, s0); s1 = fmaf(kf[7], q1.w, s1);
                s0 = fmaf(kf[8], q2.x, s0); s1 = fmaf(kf[12], q3.x, s1);
                s0 = fmaf(kf[9], q2.y, s0); s1 = fmaf(kf[13], q3.y, s1);
                s0 = fmaf(kf[10], q2.z, s0); s1 = fmaf(kf[14], q3.z, s1);
                s0 = fmaf(kf[11], q2.w, s0); s1 = fmaf(kf[15], q3.w, s1);
                float sv = s0 + s1;
#pragma unroll
                for (int of = LPP / 2; of > 0; of >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, of);
                if (sub == 0) sp[pp] = (pbase + pp < ctx) ? sv : -INFINITY;
            }
        }
        __syncwarp();
        // ---- softmax of the tile (fp32), then P.V: lane owns 4 dims
        const float v0 = sp[lane];
        const float cm = warp_max(v0);
        const float e0 = (v0 == -INFINITY) ? 0.f : __expf(v0 - cm);
        const float lsum = warp_sum(e0);
        __syncwarp();
        sp[lane] = e0;
        __syncwarp();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (lane < HD / 4) {
#pragma unroll 4
            for (int pp = 0; pp < npos; pp++) {
                float vf[4], vsc = 1.f;
                if (Q8) {
                    const uint32_t w = reinterpret_cast<const uint32_t *>(tV + pp * HD)[lane] ^ 0x80808080u;
                    vsc = __half2float(reinterpret_cast<const __half *>(tVs)[pp * NB + lane / 8]);
#pragma unroll
                    for (int k = 0; k < 4; k++) vf[k] = s8f_(w, k);
                } else {
                    const uint2 w = reinterpret_cast<const uint2 *>(tV + pp * HD * 2)[lane];
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&w.x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
                    vf[0] = f0.x; vf[1] = f0.y; vf[2] = f1.x; vf[3] = f1.y;
                }
                const float p = sp[pp] * vsc;     // the block scale rides on the probability
                a0 = fmaf(p, vf[0], a0);
                a1 = fmaf(p, vf[1], a1);
                a2 = fmaf(p, vf[2], a2);
                a3 = fmaf(p, vf[3], a3);
            }
        }
        float *o = P.partials + ((size_t)hq * P.max_tiles + tile) * (HD + 4);
        if (lane == 0) *reinterpret_cast<float2 *>(o) = make_float2(cm, lsum);
        if (lane < HD / 4) *reinterpret_cast<float4 *>(o + 4 + lane * 4) = make_float4(a0, a1, a2, a3);
        __syncwarp();
    }
}

// position in the ring as the consumer warps walk it: slot, mbarrier parity of its current fill, and the rotating
// task -> warp offset (task t of a slot goes to warp (tbase + t) % 15)
struct RingPos {
    int pos;
    uint32_t phase;
    int tbase;
};

__device__ __forceinline__ void att_task_any(const StepParams &P, const StepLayer &lw, const float *s_cos, const float *s_sin, uint32_t qflag,
                                             uint8_t *kslot, uint8_t *vslot, int UP, int tt, int kvh, int p0, int ctx, const uint8_t *xst, float *sp, int lane) {
    if (P.hd == 128) {
        if (P.kv_q8) att_task<128, true>(P, lw, s_cos, s_sin, qflag, kslot, vslot, UP, tt, kvh, p0, ctx, xst, sp, lane);
        else att_task<128, false>(P, lw, s_cos, s_sin, qflag, kslot, vslot, UP, tt, kvh, p0, ctx, xst, sp, lane);
    } else {
        if (P.kv_q8) att_task<64, true>(P, lw, s_cos, s_sin, qflag, kslot, vslot, UP, tt, kvh, p0, ctx, xst, sp, lane);
        else att_task<64, false>(P, lw, s_cos, s_sin, qflag, kslot, vslot, UP, tt, kvh, p0, ctx, xst, sp, lane);
    }
}

// log-sum-exp merge of the tiles' partials: q head hq -> attn_out[hq][hd] (half, as the reference's attention output
// tensor). Thread = (4 dims, tile group): hd/4 dim quads x 480/(hd/4) tile groups; a thread's tiles are all loaded
// before any is used (one L2 round trip for a 2k context), the tile groups meet in shared memory (the idle staging area).
__device__ __forceinline__ void att_merge(const StepParams &P, int ctx, float *scratch, int tid, uint32_t flag) {
    const int hd = P.hd, nq4 = hd >> 2, ngrp = STEP_CTHREADS / nq4, tg = tid / nq4, q4 = tid - tg * nq4;
    const int ntile = (ctx + 31) >> 5;
    constexpr int NB = 6;      // tiles per batch and thread
    for (int hq = (int)blockIdx.x; hq < P.heads; hq += (int)gridDim.x) {
        const float *pb = P.partials + (size_t)hq * P.max_tiles * (hd + 4);
        float M = -INFINITY, den = 0.f;
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t0 = tg; t0 < ntile; t0 += ngrp * NB) {
            float2 ml[NB];
            float4 a[NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const int t = t0 + i * ngrp;
                ml[i] = make_float2(-INFINITY, 0.f);
                a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < ntile) {
                    const float *pt = pb + (size_t)t * (hd + 4);
                    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(ml[i].x), "=f"(ml[i].y) : "l"(pt) : "memory");
                    a[i] = ld_cg_f32x4(pt + 4 + 4 * q4);
                }
            }
            float Mn = M;
#pragma unroll
            for (int i = 0; i < NB; i++) Mn = fmaxf(Mn, ml[i].x);
            const float c = (M == -INFINITY) ? 0.f : __expf(M - Mn);
            num.x *= c; num.y *= c; num.z *= c; num.w *= c; den *= c;
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const float w = (ml[i].x == -INFINITY) ? 0.f : __expf(ml[i].x - Mn);
                num.x = fmaf(w, a[i].x, num.x); num.y = fmaf(w, a[i].y, num.y);
                num.z = fmaf(w, a[i].z, num.z); num.w = fmaf(w, a[i].w, num.w);
                den = fmaf(w, ml[i].y, den);
            }
            M = Mn;
        }
        float *rowp = scratch + (size_t)tg * (hd + 4);
        if (q4 == 0) { rowp[0] = M; rowp[1] = den; }
        *reinterpret_cast<float4 *>(rowp + 4 + 4 * q4) = num;
        named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
        if (tid < hd) {
            float MM = -INFINITY;
            for (int g = 0; g < ngrp; g++) MM = fmaxf(MM, scratch[(size_t)g * (hd + 4)]);
            float n2 = 0.f, d2 = 0.f;
            for (int g = 0; g < ngrp; g++) {
                const float mg = scratch[(size_t)g * (hd + 4)];
                const float w = (mg == -INFINITY) ? 0.f : __expf(mg - MM);
                n2 = fmaf(w, scratch[(size_t)g * (hd + 4) + 4 + tid], n2);
                d2 = fmaf(w, scratch[(size_t)g * (hd + 4) + 1], d2);
            }
            // out as {half2, flag} words: the wo staging of every CTA polls them (no second completion count)
            const float mine = n2 / d2;
            const float nxt = __shfl_down_sync(0xffffffffu, mine, 1);
            if ((tid & 1) == 0) {
                const __half2 hv = __floats2half2_rn(mine, nxt);
                st_relaxed_sys_v2(P.attw + (((size_t)hq * hd + tid) >> 1), make_uint2(*reinterpret_cast<const uint32_t *>(&hv), flag));
            }
        }
        named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
    }
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
template <class Fmt>
__device__ __forceinline__ void stage_gemv_x(const StepParams &P, const StepOpShape &s, const void *src, bool ldcg, const __half *norm_w,
                                             uint8_t *xst, float *red, int tid) {
    const XStage a = {s.cx, nullptr, norm_w, P.eps, 0.f};
    if constexpr (is_f16_fmt<Fmt>::value) {
        if (ldcg) stage_x_f16<1, STEP_CWARPS>(a, src, reinterpret_cast<__half *>(xst), red, tid);
        else stage_x_f16<0, STEP_CWARPS>(a, src, reinterpret_cast<__half *>(xst), red, tid);
    } else {
        uint8_t *xlo = xst, *xhi = xst + (size_t)s.nxb * 16;
        float2 *xs = reinterpret_cast<float2 *>(xst + (size_t)s.nxb * 32);
        stage_x_q8_compact<Fmt::XSPLIT, STEP_CWARPS>(a, src, xlo, xhi, xs, red, tid, s.vec4 != 0, ldcg);
    }
}

// two rows (row, row + stride) against the staged activation
template <class Fmt>
__device__ __forceinline__ void dot2(const StepOpShape &s, const uint8_t *row, size_t stride, const uint8_t *xst, int lane, float *acc) {
    if constexpr (is_f16_fmt<Fmt>::value) {
        acc[0] = row_dot_f16(row, s.cx, reinterpret_cast<const __half *>(xst), lane);
        acc[1] = row_dot_f16(row + stride, s.cx, reinterpret_cast<const __half *>(xst), lane);
    } else {
        const uint4 *xlo = reinterpret_cast<const uint4 *>(xst), *xhi = reinterpret_cast<const uint4 *>(xst + (size_t)s.nxb * 16);
        const float2 *xs = reinterpret_cast<const float2 *>(xst + (size_t)s.nxb * 32);
        if constexpr (is_q4b32_fmt<Fmt>::value) {
            if (s.vec4) { row_dot_q4b32_vec4<2>(row, stride, s.nblk, xlo, xhi, xs, lane, acc); return; }
        }
        acc[0] = row_dot_q8<Fmt>(row, s.nblk, xlo, xhi, xs, lane);
        acc[1] = row_dot_q8<Fmt>(row + stride, s.nblk, xlo, xhi, xs, lane);
    }
}

template <class Fmt, bool PRIV>
__global__ void __launch_bounds__(STEP_THREADS, 1) decode_step_kernel(const __grid_constant__ StepParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[STEP_MAX_BARS], empty_bar[STEP_MAX_SLOTS];
    __shared__ int2 slot_meta[STEP_MAX_SLOTS];   // (op, unit) of what sits in each ring slot
    __shared__ float red[STEP_CWARPS];
    __shared__ unsigned long long s_key[STEP_CWARPS];
    __shared__ int s_prof[4];   // timeline probe (IFB_TRACE=1): warp 0's cycles waiting for data / in tasks, the producer's waiting for a free slot / its units

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *ring = smem;
    uint8_t *xst = smem + P.x_off;
    uint4 *resid = reinterpret_cast<uint4 *>(smem + P.resid_off);
    float *s_cos = reinterpret_cast<float *>(smem + P.rope_off), *s_sin = s_cos + P.hd / 2;
    // the per-layer pointer table lives in shared memory: with the whole SM carved out as shared memory there is next
    // to no L1, and a table lookup in global memory costs a full L2 round trip (~1 us while the weight stream runs) --
    // measured: the producer's stream was bound by exactly that, every op ran at unit_bytes / 1 us
    const StepLayer *s_layers = reinterpret_cast<const StepLayer *>(smem + P.lay_off);
    uint4 *s_norm = reinterpret_cast<uint4 *>(smem + P.norm_off);
    for (int i = tid; i < P.n_layers * (int)(sizeof(StepLayer) / 16); i += STEP_THREADS)
        reinterpret_cast<uint4 *>(smem + P.lay_off)[i] = reinterpret_cast<const uint4 *>(P.layers)[i];
    const int n_ops = P.n_ops, head_op = n_ops - 1;

    if (tid < 4) s_prof[tid] = 0;
    if (tid == 0) {
        if constexpr (PRIV) {
            for (int i = 0; i < 2 * STEP_CWARPS; i++) mbar_init(&full_bar[i], 1);
        } else {
            for (int i = 0; i < P.nslots; i++) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], STEP_CWARPS);
            }
        }
        mbar_fence_init();
    }
    // position / token / sequence number were written before this launch and only change in its last instruction
    int pos_now = P.scalars[0];
    if (pos_now < 0 || pos_now >= P.max_ctx) {      // the host checks this too (ifb_model_decode): never index past the cache
        if (tid == 0 && blockIdx.x == 0 && P.err != nullptr) { P.err[1] = pos_now; P.err[0] = STEP_ERR_CTX; }
        pos_now = max(0, min(pos_now, P.max_ctx - 1));
    }
    const int ctx = pos_now + 1;
    const uint32_t seq = (uint32_t)P.scalars[3];
    if (tid < P.hd / 2) {
        // angle_i = pos * theta_scale^i, theta_scale = theta^(-2/dims) (unary_tensor_opr.h:666-679,706-719)
        const float theta_scale = powf(P.rope_theta, -2.0f / (float)P.hd);
        float ang = (float)pos_now;
        if (tid > 0) ang *= powf(theta_scale, (float)tid);
        s_cos[tid] = cosf(ang);
        s_sin[tid] = sinf(ang);
    }
    __syncthreads();

    const int att_units = (P.parts & STEP_PART_ATT) ? P.kv_heads * ((ctx + P.att_pos - 1) / P.att_pos) : 0;
    const int tiles_per_head = (ctx + P.att_pos - 1) / P.att_pos;

    if (warp == STEP_CWARPS) {
        if constexpr (PRIV) return;        // warp-private pipelines: every consumer warp issues its own copies
        // ===== producer: depends on nothing but free ring slots =====
        if (lane != 0) return;
        const uint64_t pol = l2_policy_evict_first();   // weights and cache tiles are read once per token
        int op = 0, pos = 0;
        uint32_t phase = 1;                              // a fresh mbarrier passes a wait on parity 1
        // Units are handed out by per-op global cursors (SMs do not get equal shares of the HBM bandwidth). One claim is
        // a global atomic round trip (~1 us while the memory system is busy): claiming unit by unit caps a CTA's stream
        // at unit_bytes / latency (measured: every op ran at ~3 TB/s). So a claim takes a BATCH of consecutive units,
        // sized by what is left (guided self-scheduling: remaining / (2 * grid), at most claim_batch, down to 1 at
        // the tail), and the claim of the next batch is in flight while this one is copied.
        auto batch_of = [&](int units, int from) {
            const int b = (units - from) / (2 * (int)gridDim.x);
            return max(1, min(b, P.claim_batch));
        };
        auto units_of = [&](int o) {
            const int kind = o == head_op ? OPK_HEAD : o % 5;
            return kind == OPK_ATT ? att_units : ((P.parts & STEP_PART_GEMV) ? P.op[kind].units : 0);
        };
        unsigned issued = 0;                             // copies (ring fills) issued so far
        const uint32_t pfull0 = smem_u32(full_bar), pempty0 = smem_u32(empty_bar), ring0 = smem_u32(ring);
        const uint8_t *src0 = nullptr, *src1 = nullptr, *src2 = nullptr;   // this op's matrices / cache planes
        auto bind_op = [&](int o) {
            const int kind = o == head_op ? OPK_HEAD : o % 5;
            if (kind == OPK_HEAD) { src0 = P.lm_head; return; }
            const StepLayer &lw = s_layers[o / 5];
            if (kind == OPK_QKV) { src0 = lw.wq; src1 = lw.wk; src2 = lw.wv; }
            else if (kind == OPK_ATT) { src0 = lw.kcache; src1 = lw.vcache; }
            else if (kind == OPK_WO) src0 = lw.wo;
            else if (kind == OPK_GLU) { src0 = lw.w1; src1 = lw.w3; }
            else src0 = lw.w2;
        };
        bind_op(0);
        int units = units_of(0);
        int bsz = batch_of(units, 0);
        int bcur = atomicAdd(&P.cursors[0], bsz), bend = bcur + bsz;
        int nsz = 0, nstart = 0;
        bool have_next = false;
        while (op < n_ops) {
            const int kind = op == head_op ? OPK_HEAD : op % 5;
            const StepOpShape &s = P.op[kind];
            if (bcur >= units || bcur >= bend) {
                if (bcur < units && have_next) {         // this batch is copied: take the one requested behind it
                    bcur = nstart; bend = nstart + nsz; have_next = false;
                    continue;
                }
                // the op is handed out (a claim at or past `units`, or a batch that reached the end): next op
                if (P.trace != nullptr) {
                    trace_add(P, op, 6, (unsigned long long)s_prof[2]);
                    trace_add(P, op, 7, (unsigned long long)s_prof[3]);
                    s_prof[2] = s_prof[3] = 0;
                }
                if (++op < n_ops) {
                    bind_op(op);
                    units = units_of(op);
                    bsz = batch_of(units, 0);
                    bcur = atomicAdd(&P.cursors[op], bsz); bend = bcur + bsz;
                    have_next = false;
                }
                continue;
            }
            if (!have_next && bend < units) {            // request the following batch now: in flight while we copy this one
                nsz = batch_of(units, bend);
                nstart = atomicAdd(&P.cursors[op], nsz);
                have_next = true;
            }
            const int u = bcur++;
            // bounded run-ahead: at most max_inflight copies outstanding per SM. More only queues in the memory system
            // and every demand load of the staging code (activations, norm weights, completion counts) waits behind it.
            if (P.max_inflight > 0 && issued >= (unsigned)P.max_inflight) {
                const unsigned w = issued - (unsigned)P.max_inflight;
                mbar_wait_bounded(pfull0 + 8u * (w % (unsigned)P.nslots), (w / (unsigned)P.nslots) & 1u, P, op);
            }
            issued++;
            if (P.trace != nullptr) {
                const long long t0 = clk();
                mbar_wait_bounded(pempty0 + 8u * (uint32_t)pos, phase, P, op, 200u);
                s_prof[2] += (int)(clk() - t0);
                s_prof[3]++;
            } else
            mbar_wait_bounded(pempty0 + 8u * (uint32_t)pos, phase, P, op, 200u);
            slot_meta[pos] = make_int2(op, u);
            const uint32_t dst = ring0 + (uint32_t)pos * (uint32_t)P.slot_bytes, fb = pfull0 + 8u * (uint32_t)pos;
            const uint32_t ub = (uint32_t)s.unit_bytes;
            if (kind == OPK_ATT) {
                const int h = u / tiles_per_head, t = u - h * tiles_per_head;
                const int64_t row0 = (int64_t)h * P.max_ctx_pad + (int64_t)t * P.att_pos;
                const int hd = P.hd, nb = hd / 32, UP = P.att_pos;
                if (P.kv_q8) {
                    const uint32_t kb = (uint32_t)(UP * hd), sb = (uint32_t)(UP * nb * 2);
                    const int64_t soff = (int64_t)P.kv_heads * P.max_ctx_pad * hd;
                    mbar_arrive_expect_tx_a(fb, 2 * (kb + sb));
                    bulk_g2s_hint_a(dst, src0 + row0 * hd, kb, fb, pol);
                    bulk_g2s_hint_a(dst + kb, src0 + soff + row0 * nb * 2, sb, fb, pol);
                    bulk_g2s_hint_a(dst + kb + sb, src1 + row0 * hd, kb, fb, pol);
                    bulk_g2s_hint_a(dst + 2 * kb + sb, src1 + soff + row0 * nb * 2, sb, fb, pol);
                } else {
                    const uint32_t kb = (uint32_t)(UP * hd * 2);
                    mbar_arrive_expect_tx_a(fb, 2 * kb);
                    bulk_g2s_hint_a(dst, src0 + row0 * hd * 2, kb, fb, pol);
                    bulk_g2s_hint_a(dst + kb, src1 + row0 * hd * 2, kb, fb, pol);
                }
            } else if (kind == OPK_GLU) {
                mbar_arrive_expect_tx_a(fb, 2 * ub);
                bulk_g2s_hint_a(dst, src0 + (size_t)u * ub, ub, fb, pol);
                bulk_g2s_hint_a(dst + ub, src1 + (size_t)u * ub, ub, fb, pol);
            } else {
                const uint8_t *src = src0 + (size_t)u * ub;
                if (kind == OPK_QKV && u >= s.seg1) src = u >= s.seg2 ? src2 + (size_t)(u - s.seg2) * ub : src1 + (size_t)(u - s.seg1) * ub;
                mbar_arrive_expect_tx_a(fb, ub);
                bulk_g2s_hint_a(dst, src, ub, fb, pol);
            }
            if (++pos == P.nslots) { pos = 0; phase ^= 1u; }
        }
        mbar_wait_bounded(pempty0 + 8u * (uint32_t)pos, phase, P, n_ops);
        slot_meta[pos] = make_int2(n_ops, 0);            // end of the token
        mbar_arrive_a(pfull0 + 8u * (uint32_t)pos);
        return;
    }

    // ===== consumers =====
    float *sp = reinterpret_cast<float *>(smem + P.sp_off) + (size_t)warp * 32;
    const int tp = P.tp_size;
    // The outputs of the row-parallel GEMVs (wo, w2) travel as flag-in-data words {half2, flag}: under tensor parallelism
    // into every rank's exchange region, on one GPU into this GPU's own (local_ll). The reader polls the words it needs
    // -- no completion count, fence or atomic on that path, one L2 round trip instead of three.
    const bool ll = tp > 1 || P.local_ll != 0;

    // staging of op `op`'s input (every CTA runs it for every op, in order: the residual stream lives in each CTA's
    // shared memory, and "this CTA has staged op k" is part of what its completion count of op k tells the peers)
    auto stage_for = [&](int op) {
        const int kind = op == head_op ? OPK_HEAD : op % 5, layer = op / 5;
        trace_op(P, op, 0, tid);
        const uint32_t lflag = (seq << 8) | (uint32_t)(layer + 1);      // tag of this layer's qkv / attention-output words
        if (kind == OPK_ATT) {
            // q of every head: polled from the words the wq|wk|wv GEMV wrote (no completion count on this path)
            trace_op(P, op, 1, tid);
            if (P.kv_q8) stage_att<true>(P, xst, s_cos, s_sin, tid, lflag, op);
            else stage_att<false>(P, xst, s_cos, s_sin, tid, lflag, op);
        } else {
            if (kind == OPK_WO) {
                dep_wait(&P.done[op - 1], gridDim.x, P, op, tid);      // every tile's partial is stored
                att_merge(P, ctx, reinterpret_cast<float *>(xst), tid, lflag);
            }
            // where the 8 values of group g come from (and, for the ops that start from the residual stream, the
            // residual update itself: x_new = resid + row-parallel output, stored back to the shared-memory stream)
            const bool from_resid = kind == OPK_QKV || kind == OPK_GLU || kind == OPK_HEAD;
            const uint32_t xidx = kind == OPK_GLU ? (uint32_t)(2 * layer + 1) : (uint32_t)(2 * (kind == OPK_HEAD ? P.n_layers : layer));
            const uint32_t xflag = (seq << 8) | xidx;
            const uint8_t *xbase = reinterpret_cast<const uint8_t *>(P.peer[P.tp_rank]) + P.x_recv + (long long)(xidx & 1u) * P.x_recv_parity;
            const bool gather = ll && (P.parts & STEP_PART_GEMV);
            if (from_resid && op != 0 && !ll) dep_wait(&P.done[op - 1], gridDim.x, P, op, tid);
            auto load = [&](int g) -> uint4 {
                if (kind == OPK_WO) return ll_load8(reinterpret_cast<const uint4 *>(P.attw), g, lflag, P, op);   // written by this step's merge in any case
                if (kind == OPK_W2) return ll_load8(reinterpret_cast<const uint4 *>(P.hbufw), g, lflag, P, op, (P.parts & STEP_PART_GEMV) != 0);
                if (op == 0) {
                    // GetEmbdTensor (inference_engine.cc:1324-1350): the token's embedding row starts the residual stream
                    int tok = P.scalars[1];
                    tok = max(0, min(tok, P.vocab - 1));
                    const uint4 v = reinterpret_cast<const uint4 *>(P.tok_emb + (size_t)tok * P.d)[g];
                    resid[g] = v;
                    return v;
                }
                float acc[8];
                if (gather) {
                    // partial sums of all ranks, added in rank order with fp16 rounding after every add (MergeTensors,
                    // inference_worker.cc:2264-2289); one GPU: the single "rank" is this GPU's own wo / w2 output
                    for (int p = 0; p < P.tp_size; p++) {
                        const uint4 d4 = ll_load8(reinterpret_cast<const uint4 *>(xbase + (size_t)p * P.x_recv_rank), g, xflag, P, op);
                        const uint32_t dw[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&dw[k]));
                            if (p == 0) { acc[2 * k] = f.x; acc[2 * k + 1] = f.y; }
                            else { acc[2 * k] = round_h(acc[2 * k] + f.x); acc[2 * k + 1] = round_h(acc[2 * k + 1] + f.y); }
                        }
                    }
                } else {
                    float t8[8];
                    unpack8(ll ? make_uint4(0, 0, 0, 0) : ld_cg128(reinterpret_cast<const uint4 *>(P.ybuf) + g), t8);
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[k] = t8[k];
                }
                // + residual, the last add (TensorOpr::Add, inference_worker.cc:849)
                const uint4 rv = resid[g];
                uint4 out;
                __half2 *o = reinterpret_cast<__half2 *>(&out);
                const __half2 *r = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 rf = __half22float2(r[i]);
                    o[i] = __floats2half2_rn(acc[2 * i] + rf.x, acc[2 * i + 1] + rf.y);
                }
                resid[g] = out;
                return out;
            };
            const __half *nw = !from_resid ? nullptr : (kind == OPK_HEAD ? P.out_norm : (kind == OPK_QKV ? s_layers[layer].attn_norm : s_layers[layer].ffn_norm));
            bool fp16_x = kind == OPK_HEAD && P.head_f16;
            if constexpr (is_f16_fmt<Fmt>::value) fp16_x = true;
            if (fp16_x) {
                // fp16 activation (fp16 weights / fp16 lm_head): without a norm the staged form IS the vector; with one,
                // the residual stream (just updated in shared memory) goes through the fp16 staging
                const int ngx = P.op[kind].cx >> 3;
                for (int g = tid; g < ngx; g += STEP_CTHREADS) {
                    const uint4 v = load(g);
                    if (!from_resid) reinterpret_cast<uint4 *>(xst)[g] = v;
                }
                trace_op(P, op, 1, tid);
                if (from_resid) {
                    named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
                    const XStage a = {P.op[kind].cx, nullptr, nw, P.eps, 0.f};
                    stage_x_f16<0, STEP_CWARPS>(a, resid, reinterpret_cast<__half *>(xst), red, tid);
                }
            } else {
                if constexpr (!is_f16_fmt<Fmt>::value) stage_fused_q8<Fmt::XSPLIT>(P.op[kind], nw, P.eps, xst, red, tid, load);
                trace_op(P, op, 1, tid);
            }
        }
        named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
        trace_op(P, op, 2, tid);
    };

    // The slot walk. Every consumer warp passes every slot (it arrives on the slot's "empty" barrier whether it had a
    // task in it or not), so what a pass costs is paid 15 x slots-per-token times: everything that is constant for an
    // op is computed once per op, and each op kind has its own tight slot loop (ncu on the first version: dp4a was 4 % of
    // the executed instructions, the generic per-slot bookkeeping ~45 %).
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    if constexpr (PRIV) {
        // ===== warp-private pipelines =====
        // Every consumer warp streams its OWN tasks through its OWN two slots: lane 0 issues the bulk copies, the warp
        // waits on its own two barriers -- no producer warp, no unit cursors, no slot every warp has to pass. Tasks are
        // dealt statically, round-robin over all grid x 15 warps and continuing across ops (task t of an op whose first
        // task fell to global warp `rot` belongs to warp (rot + t) mod nwt), so every warp streams the same number of
        // chunks +-1 per op and the ops of a layer even each other out. The copy side runs two chunks ahead of the
        // arithmetic and does not stop at op boundaries: weights and cache tiles depend on no activation, so while the
        // CTAs synchronise at a boundary the next op's first chunks are already on their way.
        const int nwt = STEP_CWARPS * (int)gridDim.x;
        const int gw = warp * (int)gridDim.x + (int)blockIdx.x;
        const uint64_t pol = l2_policy_evict_first();
        const uint32_t bar0 = smem_u32(full_bar) + 8u * (uint32_t)(2 * warp), ring0 = smem_u32(ring) + (uint32_t)(2 * warp) * (uint32_t)P.pslot;
        uint8_t *myring = ring + (size_t)(2 * warp) * P.pslot;
        const int pslot = P.pslot;
        const bool gemv_on = (P.parts & STEP_PART_GEMV) != 0;
        auto kind_of = [&](int o) { return o == head_op ? OPK_HEAD : o % 5; };
        auto ntasks_of = [&](int kind) { return kind == OPK_ATT ? att_units : (gemv_on ? P.pv[kind].ntasks : 0); };
        auto nchunks_of = [&](int kind) { return kind == OPK_ATT ? 2 : (kind == OPK_GLU ? 2 * P.pv[kind].nc : P.pv[kind].nc); };

        // ---- the copy side: a cursor over this warp's chunks of the whole token
        int i_op = 0, i_kind = OPK_QKV, i_rot = 0, i_nt = ntasks_of(OPK_QKV), i_nch = nchunks_of(OPK_QKV), i_t = gw, i_c = 0;
        unsigned i_seq = 0;
        auto issue_one = [&]() {
            while (i_op < n_ops && i_t >= i_nt) {            // this warp has no (more) tasks in the op: next op
                i_rot += i_nt % nwt; if (i_rot >= nwt) i_rot -= nwt;
                if (++i_op >= n_ops) break;
                i_kind = kind_of(i_op); i_nt = ntasks_of(i_kind); i_nch = nchunks_of(i_kind);
                i_t = gw - i_rot; if (i_t < 0) i_t += nwt;
                i_c = 0;
            }
            if (i_op >= n_ops) return;
            const uint32_t sl = i_seq & 1u, fb = bar0 + 8u * sl, dst = ring0 + sl * (uint32_t)pslot;
            if (lane == 0) {
                const StepLayer &lw = s_layers[i_kind == OPK_HEAD ? 0 : i_op / 5];
                if (i_kind == OPK_ATT) {
                    const int h = i_t / tiles_per_head, tl = i_t - h * tiles_per_head;
                    const int64_t row0 = (int64_t)h * P.max_ctx_pad + (int64_t)tl * 32;
                    const int hd = P.hd, nb = hd >> 5;
                    const uint8_t *plane = i_c == 0 ? lw.kcache : lw.vcache;
                    if (P.kv_q8) {
                        const uint32_t kb = (uint32_t)(32 * hd), sb = (uint32_t)(32 * nb * 2);
                        const int64_t soff = (int64_t)P.kv_heads * P.max_ctx_pad * hd;
                        mbar_arrive_expect_tx_a(fb, kb + sb);
                        bulk_g2s_hint_a(dst, plane + row0 * hd, kb, fb, pol);
                        bulk_g2s_hint_a(dst + kb, plane + soff + row0 * nb * 2, sb, fb, pol);
                    } else {
                        const uint32_t kb = (uint32_t)(32 * hd * 2);
                        mbar_arrive_expect_tx_a(fb, kb);
                        bulk_g2s_hint_a(dst, plane + row0 * hd * 2, kb, fb, pol);
                    }
                } else {
                    const StepPrivShape &v = P.pv[i_kind];
                    const int rb = P.op[i_kind].row_bytes;
                    int r = i_t * v.rt, cc = i_c;
                    const uint8_t *base;
                    if (i_kind == OPK_QKV) {
                        if (r < P.q_dim) base = lw.wq;
                        else if (r < P.q_dim + P.kv_dim) { base = lw.wk; r -= P.q_dim; }
                        else { base = lw.wv; r -= P.q_dim + P.kv_dim; }
                    } else if (i_kind == OPK_WO) base = lw.wo;
                    else if (i_kind == OPK_GLU) { base = cc >= v.nc ? lw.w3 : lw.w1; if (cc >= v.nc) cc -= v.nc; }
                    else if (i_kind == OPK_W2) base = lw.w2;
                    else base = P.lm_head;
                    const uint8_t *src = base + (size_t)r * rb;
                    if (v.nc == 1) {
                        const uint32_t nb_ = (uint32_t)(v.rt * rb);
                        mbar_arrive_expect_tx_a(fb, nb_);
                        bulk_g2s_hint_a(dst, src, nb_, fb, pol);
                    } else {                                    // rt == 2: the same K-segment of two rows
                        const uint32_t off = (uint32_t)(cc * v.kc_bytes), seg = min((uint32_t)v.kc_bytes, (uint32_t)rb - off);
                        mbar_arrive_expect_tx_a(fb, 2 * seg);
                        bulk_g2s_hint_a(dst, src + off, seg, fb, pol);
                        bulk_g2s_hint_a(dst + seg, src + rb + off, seg, fb, pol);
                    }
                }
            }
            i_seq++;
            if (++i_c == i_nch) { i_c = 0; i_t += nwt; }
        };
        issue_one();
        issue_one();

        // ---- the arithmetic side
        unsigned c_seq = 0;
        // wait for this warp's chunk c_seq; returns its slot
        auto chunk_wait = [&](int op) -> uint8_t * {
            const uint32_t sl = c_seq & 1u;
            mbar_wait_bounded(bar0 + 8u * sl, (c_seq >> 1) & 1u, P, op);
            return myring + (size_t)sl * pslot;
        };
        // the chunk is read: its slot takes the next copy (the warp's own generic-proxy reads, then an async-proxy write)
        auto chunk_done = [&]() {
            c_seq++;
            __syncwarp();
            if (lane == 0) fence_proxy_async_smem();
            issue_one();
        };
        int rot = 0;
        for (int op = 0; op < n_ops; op++) {
            const int kind = kind_of(op), layer = op / 5;
            if (op > 0) {
                // end of op - 1: the count where somebody waits for one (attention tiles -> merge, lm_head -> end of the
                // token, row-parallel outputs that do not travel as flag-in-data words); a CTA barrier otherwise (the
                // staging below overwrites what the slowest warp may still read)
                const int k = op - 1, kk = kind_of(k);
                if (kk == OPK_ATT || (!ll && (kk == OPK_WO || kk == OPK_W2))) dep_signal(&P.done[k], tid);
                else named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
            }
            stage_for(op);
            trace_op(P, op, 3, tid);
            const StepOpShape &s = P.op[kind];
            const int nt = ntasks_of(kind);
            const bool compute = (P.debug & 1) == 0;
            const uint32_t lflag_op = (seq << 8) | (uint32_t)(layer + 1);
            int t = gw - rot; if (t < 0) t += nwt;
            rot += nt % nwt; if (rot >= nwt) rot -= nwt;
            if (kind == OPK_ATT) {
                for (; t < nt; t += nwt) {
                    // the K tile and the V tile of the task are this warp's chunks c_seq and c_seq + 1: both slots at once
                    const unsigned n1 = c_seq + 1u;
                    mbar_wait_bounded(bar0 + 8u * (c_seq & 1u), (c_seq >> 1) & 1u, P, op);
                    mbar_wait_bounded(bar0 + 8u * (n1 & 1u), (n1 >> 1) & 1u, P, op);
                    uint8_t *kslot = myring + (size_t)(c_seq & 1u) * pslot, *vslot = myring + (size_t)(n1 & 1u) * pslot;
                    const int h = t / tiles_per_head, tl = t - h * tiles_per_head;
                    if (compute) att_task_any(P, s_layers[layer], s_cos, s_sin, lflag_op, kslot, vslot, 32, 0, h, tl * 32, ctx, xst, sp, lane);
                    c_seq += 2;
                    __syncwarp();
                    if (lane == 0) fence_proxy_async_smem();
                    issue_one();
                    issue_one();
                }
                continue;
            }
            const StepPrivShape &v = P.pv[kind];
            const int rt = v.rt, nc = v.nc, np = rt >> 1;
            const bool f16w = is_f16_fmt<Fmt>::value || (kind == OPK_HEAD && P.head_f16 != 0);
            // one chunk of rt rows (stride = the chunk's bytes per row) against the staged activation, sums continued in acc
            auto chunk_dot = [&](const uint8_t *slot, int cc, float *acc) {
                if (f16w) {
                    const int e0 = cc * v.kc_blk, ne = nc == 1 ? s.cx : min(v.kc_blk, s.cx - e0);
                    const size_t stride = (size_t)ne * 2;
                    const __half *xh = reinterpret_cast<const __half *>(xst) + e0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (j < rt) acc[j] = row_dot_f16(slot + (size_t)j * stride, ne, xh, lane, acc[j]);
                    return;
                }
                if constexpr (!is_f16_fmt<Fmt>::value) {
                    const int b0 = cc * v.kc_blk, nb_ = nc == 1 ? s.nblk : min(v.kc_blk, s.nblk - b0);
                    const size_t stride = (size_t)nb_ * Fmt::BYTES;
                    const int xo = b0 * (Fmt::CAP / 32);
                    const uint4 *xlo = reinterpret_cast<const uint4 *>(xst) + xo, *xhi = reinterpret_cast<const uint4 *>(xst + (size_t)s.nxb * 16) + xo;
                    const float2 *xs = reinterpret_cast<const float2 *>(xst + (size_t)s.nxb * 32) + xo;
                    if constexpr (is_q4b32_fmt<Fmt>::value) {
                        if (s.vec4) {
#pragma unroll
                            for (int p = 0; p < 2; p++)
                                if (p < np) row_dot_q4b32_vec4<2, false>(slot + (size_t)(2 * p) * stride, stride, nb_, xlo, xhi, xs, lane, acc + 2 * p);
                            return;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (j < rt) acc[j] = row_dot_q8<Fmt>(slot + (size_t)j * stride, nb_, xlo, xhi, xs, lane, acc[j]);
                }
            };
            if (kind == OPK_GLU) {
                for (; t < nt; t += nwt) {
                    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a3[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int cc = 0; cc < nc; cc++) {
                        const uint8_t *slot = chunk_wait(op);
                        if (compute) chunk_dot(slot, cc, a1);
                        chunk_done();
                    }
                    for (int cc = 0; cc < nc; cc++) {
                        const uint8_t *slot = chunk_wait(op);
                        if (compute) chunk_dot(slot, cc, a3);
                        chunk_done();
                    }
                    if (!compute) continue;
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        if (p >= np) break;
                        float h2[2];
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            // SiluActivation_Kernel (unary_tensor_opr.h:553) then Mul, each stored as half
                            const float v0 = round_h(warp_sum(a1[2 * p + j])), v1 = round_h(warp_sum(a3[2 * p + j]));
                            const float g = round_h(v0 / (1.0f + expf(-v0)));
                            h2[j] = g * v1;
                        }
                        if (lane == 0) {
                            const __half2 hv = __floats2half2_rn(h2[0], h2[1]);
                            st_relaxed_sys_v2(P.hbufw + ((t * rt) >> 1) + p, make_uint2(*reinterpret_cast<const uint32_t *>(&hv), lflag_op));
                        }
                    }
                }
                continue;
            }
            const bool qkv = kind == OPK_QKV, head = kind == OPK_HEAD;
            const bool push = !qkv && !head && ll;
            uint32_t *out32 = reinterpret_cast<uint32_t *>(P.ybuf);
            const uint32_t xidx = kind == OPK_WO ? (uint32_t)(2 * layer + 1) : (uint32_t)(2 * layer + 2);
            const size_t xoff = (size_t)P.x_recv + (size_t)(xidx & 1u) * P.x_recv_parity + (size_t)P.tp_rank * P.x_recv_rank;
            for (; t < nt; t += nwt) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int cc = 0; cc < nc; cc++) {
                    const uint8_t *slot = chunk_wait(op);
                    if (compute) chunk_dot(slot, cc, acc);
                    chunk_done();
                }
                if (!compute) continue;
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    if (p >= np) break;
                    const float s0 = warp_sum(acc[2 * p]), s1 = warp_sum(acc[2 * p + 1]);
                    if (lane != 0) continue;
                    const int o = t * rt + 2 * p;
                    if (head) {
                        const float v0 = round_h(s0), v1 = round_h(s1);
                        const __half2 hv = __floats2half2_rn(v0, v1);
                        const uint32_t bits = *reinterpret_cast<const uint32_t *>(&hv);
                        if (tp == 1) {
                            reinterpret_cast<uint32_t *>(P.logits)[o >> 1] = bits;
                            if (v0 == v0 && (v0 > best_v || (v0 == best_v && o < best_i))) { best_v = v0; best_i = o; }
                            if (v1 == v1 && (v1 > best_v || (v1 == best_v && o + 1 < best_i))) { best_v = v1; best_i = o + 1; }
                        } else {
                            const uint2 wd = make_uint2(bits, (seq << 8) | 255u);
                            const size_t off = (size_t)P.x_logits + ((size_t)(P.tp_rank * P.vocab_l + o) >> 1) * 8;
                            for (int pr = 0; pr < tp; pr++) st_relaxed_sys_v2(reinterpret_cast<uint8_t *>(P.peer[pr]) + off, wd);
                        }
                        continue;
                    }
                    const __half2 hv = __floats2half2_rn(s0, s1);      // GEMV output tensors are half
                    const uint32_t bits = *reinterpret_cast<const uint32_t *>(&hv);
                    if (qkv) {
                        st_relaxed_sys_v2(P.qkvw + (o >> 1), make_uint2(bits, lflag_op));   // q|k|v as {half2, flag} words
                    } else if (!push) {
                        out32[o >> 1] = bits;
                    } else {
                        const uint2 wd = make_uint2(bits, (seq << 8) | xidx);
                        for (int pr = 0; pr < tp; pr++) st_relaxed_sys_v2(reinterpret_cast<uint8_t *>(P.peer[pr]) + xoff + (size_t)(o >> 1) * 8, wd);
                    }
                }
            }
        }
        // end of the lm_head
        if (tp == 1 && lane == 0 && best_i != 0x7fffffff) atomicMax(P.amax, amax_key(best_v, best_i));
        dep_signal(&P.done[head_op], tid);
    } else {
    int cur_op = -1;            // nothing staged yet: the first pass stages op 0 while the first units are in flight
    int next_op = 0;
    int pos = 0, tbase = 0;
    uint32_t phase = 0;
    const int nslots = P.nslots, slot_bytes = P.slot_bytes;
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    // one pass over a slot: wait for its data; false when it belongs to another op (left in place for the transition)
#define IFB_SLOT_BEGIN()                                                            \
        mbar_wait_bounded(full0 + 8u * (uint32_t)pos, phase, P, cur_op);            \
        const int2 meta = slot_meta[pos];                                           \
        if (meta.x != cur_op) { next_op = meta.x; break; }                          \
        uint8_t *slot = ring + (size_t)pos * slot_bytes;                            \
        const int u = meta.y;                                                       \
        int t = warp - tbase;                                                       \
        if (t < 0) t += STEP_CWARPS;
#define IFB_SLOT_END()                                                              \
        tbase += ntask_mod;                                                         \
        if (tbase >= STEP_CWARPS) tbase -= STEP_CWARPS;                             \
        __syncwarp();                                                               \
        if (lane == 0) mbar_arrive_a(empty0 + 8u * (uint32_t)pos);                  \
        if (++pos == nslots) { pos = 0; phase ^= 1u; }
    for (;;) {
        // ---- op boundary: signal what is done, stage what comes
        for (int k = cur_op; k < next_op; k++) {
            if (k >= 0) {
                if (k == head_op && tp == 1 && lane == 0 && best_i != 0x7fffffff) atomicMax(P.amax, amax_key(best_v, best_i));
                // "this CTA's part of op k is stored (and its staging of op k is long done)": counted where somebody waits
                // for the count -- the attention tiles (merge), the lm_head (end of the token), and the row-parallel
                // outputs when they do not travel as flag-in-data words. The q|k|v, attention-output and h words are their
                // own signal. The count of ATT is the ONE grid-wide barrier of a layer, and it is what keeps every reused
                // buffer safe: the words / partials / exchange buffers of layer l+1 are written behind it, and a CTA counts
                // itself only after all its stagings up to ATT(l+1) -- i.e. after its last read of the layer-l contents.
                const int kk = k == head_op ? OPK_HEAD : k % 5;
                if (kk == OPK_ATT || kk == OPK_HEAD || (!ll && (kk == OPK_WO || kk == OPK_W2)))
                    dep_signal(&P.done[k], tid);
                else
                    named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);   // the staging below overwrites what the slowest warp may still read
            }
            if (k + 1 < n_ops) stage_for(k + 1);
        }
        cur_op = next_op;
        if (cur_op >= n_ops) break;
        trace_op(P, cur_op, 3, tid);
        const int kind = cur_op == head_op ? OPK_HEAD : cur_op % 5, layer = cur_op / 5;
        const StepOpShape &s = P.op[kind];
        const int ntask = (P.debug & 1) ? 0 : s.tasks, ntask_mod = s.tasks % STEP_CWARPS;
        const uint32_t lflag_op = (seq << 8) | (uint32_t)(layer + 1);
        const long long t_op0 = (P.trace != nullptr && tid == 0) ? clk() : 0ll;
        if (kind == OPK_ATT) {
            for (;;) {
                IFB_SLOT_BEGIN()
                if (t < ntask) {
                    const int h = u / tiles_per_head, tl = u - h * tiles_per_head;
                    uint8_t *vslot = slot + (P.kv_q8 ? P.att_pos * (P.hd + (P.hd >> 5) * 2) : 32 * P.hd * 2);
                    for (; t < ntask; t += STEP_CWARPS) att_task_any(P, s_layers[layer], s_cos, s_sin, lflag_op, slot, vslot, P.att_pos, t, h, tl * P.att_pos, ctx, xst, sp, lane);
                }
                IFB_SLOT_END()
            }
        } else if (kind == OPK_GLU) {
            // a task = two consecutive (w1, w3) row pairs: its two outputs h[o], h[o+1] leave as ONE {half2, flag} word
            const int R = s.R, row_bytes_ = s.row_bytes;
            const size_t stride = (size_t)s.unit_bytes;
            for (;;) {
                IFB_SLOT_BEGIN()
                for (; t < ntask; t += STEP_CWARPS) {
                    float h2[2];
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        float acc[2];
                        dot2<Fmt>(s, slot + (size_t)(2 * t + j) * row_bytes_, stride, xst, lane, acc);
                        acc[0] = warp_sum(acc[0]);
                        acc[1] = warp_sum(acc[1]);
                        // SiluActivation_Kernel (unary_tensor_opr.h:553) then Mul, each stored as half
                        const float v0 = round_h(acc[0]), v1 = round_h(acc[1]);   // GEMV output tensors are half
                        const float g = round_h(v0 / (1.0f + expf(-v0)));
                        h2[j] = g * v1;
                    }
                    if (lane == 0) {
                        const __half2 hv = __floats2half2_rn(h2[0], h2[1]);
                        st_relaxed_sys_v2(P.hbufw + ((u * R + 2 * t) >> 1), make_uint2(*reinterpret_cast<const uint32_t *>(&hv), lflag_op));
                    }
                }
                IFB_SLOT_END()
            }
        } else if (kind == OPK_HEAD) {
            const int R = s.R, row_bytes_ = s.row_bytes;
            const bool f16 = P.head_f16 != 0;
            for (;;) {
                IFB_SLOT_BEGIN()
                for (; t < ntask; t += STEP_CWARPS) {
                    float acc[2];
                    const uint8_t *row = slot + (size_t)(2 * t) * row_bytes_;
                    if (f16) {
                        acc[0] = row_dot_f16(row, s.cx, reinterpret_cast<const __half *>(xst), lane);
                        acc[1] = row_dot_f16(row + row_bytes_, s.cx, reinterpret_cast<const __half *>(xst), lane);
                    } else {
                        dot2<Fmt>(s, row, (size_t)row_bytes_, xst, lane, acc);
                    }
                    acc[0] = warp_sum(acc[0]);
                    acc[1] = warp_sum(acc[1]);
                    if (lane == 0) {
                        const int o = u * R + 2 * t;
                        const float v0 = round_h(acc[0]), v1 = round_h(acc[1]);
                        const __half2 hv = __floats2half2_rn(v0, v1);
                        const uint32_t bits = *reinterpret_cast<const uint32_t *>(&hv);
                        if (tp == 1) {
                            reinterpret_cast<uint32_t *>(P.logits)[o >> 1] = bits;
                            if (v0 == v0 && (v0 > best_v || (v0 == best_v && o < best_i))) { best_v = v0; best_i = o; }
                            if (v1 == v1 && (v1 > best_v || (v1 == best_v && o + 1 < best_i))) { best_v = v1; best_i = o + 1; }
                        } else {
                            const uint2 wd = make_uint2(bits, (seq << 8) | 255u);
                            const size_t off = (size_t)P.x_logits + ((size_t)(P.tp_rank * P.vocab_l + o) >> 1) * 8;
                            for (int p = 0; p < tp; p++) st_relaxed_sys_v2(reinterpret_cast<uint8_t *>(P.peer[p]) + off, wd);
                        }
                    }
                }
                IFB_SLOT_END()
            }
        } else {   // QKV, WO, W2: two adjacent rows per task, plain half outputs
            const int R = s.R, row_bytes_ = s.row_bytes;
            uint32_t *out32 = reinterpret_cast<uint32_t *>(P.ybuf);
            const bool qkv = kind == OPK_QKV;
            const bool push = !qkv && ll;
            // row-parallel partial sums go straight into every rank's exchange region over NVLink
            const uint32_t xidx = kind == OPK_WO ? (uint32_t)(2 * layer + 1) : (uint32_t)(2 * layer + 2);
            const size_t xoff = (size_t)P.x_recv + (size_t)(xidx & 1u) * P.x_recv_parity + (size_t)P.tp_rank * P.x_recv_rank;
            for (;;) {
                IFB_SLOT_BEGIN()
                for (; t < ntask; t += STEP_CWARPS) {
                    float acc[2];
                    dot2<Fmt>(s, slot + (size_t)(2 * t) * row_bytes_, (size_t)row_bytes_, xst, lane, acc);
                    acc[0] = warp_sum(acc[0]);
                    acc[1] = warp_sum(acc[1]);
                    if (lane == 0) {
                        const int o = u * R + 2 * t;
                        const __half2 hv = __floats2half2_rn(acc[0], acc[1]);      // GEMV output tensors are half
                        const uint32_t bits = *reinterpret_cast<const uint32_t *>(&hv);
                        if (qkv) {
                            st_relaxed_sys_v2(P.qkvw + (o >> 1), make_uint2(bits, lflag_op));   // q|k|v as {half2, flag} words
                        } else if (!push) {
                            out32[o >> 1] = bits;
                        } else {
                            const uint2 wd = make_uint2(bits, (seq << 8) | xidx);
                            for (int p = 0; p < tp; p++) st_relaxed_sys_v2(reinterpret_cast<uint8_t *>(P.peer[p]) + xoff + (size_t)(o >> 1) * 8, wd);
                        }
                    }
                }
                IFB_SLOT_END()
            }
        }
        if (P.trace != nullptr && tid == 0) trace_add(P, cur_op, 5, (unsigned long long)(clk() - t_op0));   // slot walk of this op, warp 0
    }
#undef IFB_SLOT_BEGIN
#undef IFB_SLOT_END
    }   // shared ring

    // ===== end of the token: next token id, position += 1 =====
    if (tp > 1 && (P.parts & STEP_PART_GEMV)) {
        // the vocabulary-sharded lm_head of every rank pushed its logits here as flag-in-data words: each CTA
        // compacts a slice to fp16 and folds it into the argmax key
        const uint2 *ll = reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(P.peer[P.tp_rank]) + P.x_logits);
        const uint32_t flag = (seq << 8) | 255u;
        const int nw = P.vocab >> 1;
        const int per = (nw + (int)gridDim.x - 1) / (int)gridDim.x;
        const int w0 = (int)blockIdx.x * per, w1 = min(nw, w0 + per);
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int w = w0 + tid; w < w1; w += STEP_CTHREADS) {
            uint2 v = ld_relaxed_sys_v2(ll + w);
            if (v.y != flag) {
                const unsigned long long t0 = globaltimer_ns();
                unsigned it = 0;
                while (v.y != flag) {
                    v = ld_relaxed_sys_v2(ll + w);
                    if ((++it & 255u) == 0u && globaltimer_ns() - t0 > P.timeout_ns) step_abort(P, STEP_ERR_LOGITS, head_op, w);
                }
            }
            reinterpret_cast<uint32_t *>(P.logits)[w] = v.x;
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
            if (f.x == f.x && f.x > bv) { bv = f.x; bi = 2 * w; }
            if (f.y == f.y && f.y > bv) { bv = f.y; bi = 2 * w + 1; }
        }
        unsigned long long key = bi != 0x7fffffff ? amax_key(bv, bi) : 0ull;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, of);
            key = ok > key ? ok : key;
        }
        if (lane == 0) s_key[warp] = key;
        named_bar_sync(BAR_CONSUMERS, STEP_CTHREADS);
        if (tid == 0) {
            for (int i = 1; i < STEP_CWARPS; i++) key = s_key[i] > key ? s_key[i] : key;
            if (key != 0ull) atomicMax(P.amax, key);
        }
        dep_signal(&P.done[n_ops + P.n_layers], tid);
    }
    if (blockIdx.x == 0) {
        dep_wait(&P.done[(tp > 1 && (P.parts & STEP_PART_GEMV)) ? n_ops + P.n_layers : head_op], gridDim.x, P, n_ops, tid);
        for (int g = tid; g < (P.d >> 3); g += STEP_CTHREADS) reinterpret_cast<uint4 *>(P.x_out)[g] = resid[g];
        if (tid == 0) {
            const unsigned long long key = ld_cg_u64(P.amax);
            const int tok = key == 0ull ? 0 : (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
            P.scalars[2] = tok;
            if (P.advance) { P.scalars[0] = pos_now + 1; P.scalars[1] = tok; }
            P.scalars[3] = (int32_t)(seq + 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool ax8_fmt(int dt) {
    switch (dt) {
        case DT_Q8_B32T2: case DT_Q6_B64T1: case DT_Q5_B64T1: case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: case DT_Q4_B64T1: case DT_Q3H_B64T1: return true;
        default: return false;
    }
}

int64_t step_counter_bytes(const StepParams &P) {
    return ((int64_t)P.n_ops * 4 + ((int64_t)P.n_ops + P.n_layers + 1) * 4 + 16 + 255) / 256 * 256;
}

// rows per unit: R * row_bytes * nmat <= target, R divides every segment, R*row_bytes % 16 == 0 (bulk-copy
// granularity), R even for ops whose task is two adjacent rows; not larger than rows/grid when that leaves >= 2
static int pick_rows(const int *seg_rows, int nseg, int row_bytes_, int nmat, bool even, int grid, int target) {
    int total = 0;
    for (int i = 0; i < nseg; i++) total += seg_rows[i];
    const int byte_cap = target / (row_bytes_ * nmat);
    const int fair = total / (4 * grid);       // aim for >= 4 units per CTA: a short op is all tail otherwise
    for (int pass = 0; pass < 2; pass++) {
        int cap = byte_cap;
        if (pass == 0) {                       // first choice: at least one unit per CTA
            if (fair < 2 || cap <= fair) continue;
            cap = fair;
        }
        for (int r = cap; r >= 1; r--) {
            if (even && (r & 1)) continue;
            if (((int64_t)r * row_bytes_) % 16 != 0) continue;
            bool ok = true;
            for (int i = 0; i < nseg; i++) ok = ok && seg_rows[i] % r == 0;
            if (ok) return r;
        }
    }
    return 0;
}

int step_build(StepParams &P, int kv_dtype, int head_dtype, int sms) {
    const int wdt = P.w_dtype;
    IFB_CHECK_ARG(wdt == DT_F16 || ax8_fmt(wdt), "step kernel: weight type %d has no dp4a path", wdt);
    IFB_CHECK_ARG(head_dtype == DT_F16 || head_dtype == wdt, "step kernel: lm_head type %d differs from the weight type %d", head_dtype, wdt);
    IFB_CHECK_ARG(P.hd == 64 || P.hd == 128, "step kernel: head_dim %d (64 / 128)", P.hd);
    IFB_CHECK_ARG(P.group == 1 || P.group == 2 || P.group == 4 || P.group == 8, "step kernel: heads/kv_heads = %d", P.group);
    IFB_CHECK_ARG(kv_dtype == DT_F16 || kv_dtype == DT_Q8_B32T2, "step kernel: kv cache type %d", kv_dtype);
    IFB_CHECK_ARG(P.d % 8 == 0 && P.q_dim % 8 == 0 && P.ffn % 8 == 0 && P.vocab_l % 2 == 0 && P.vocab % 2 == 0, "step kernel: dims must be multiples of 8");
    P.kv_q8 = kv_dtype == DT_Q8_B32T2 ? 1 : 0;
    P.head_f16 = head_dtype == DT_F16 ? 1 : 0;
    P.att_pos = P.kv_q8 ? 64 : 32;
    P.n_ops = 5 * P.n_layers + 1;
    P.max_tiles = (P.max_ctx + 31) / 32;
    const int cap = wdt == DT_F16 ? 32 : block_capacity(wdt);

    auto gemv_shape = [&](int kind, const int *segs, int nseg, int cx, int dt, int nmat) -> int {
        StepOpShape &s = P.op[kind];
        s = StepOpShape();
        s.cx = cx;
        s.row_bytes = (int)(dt == DT_F16 ? (int64_t)cx * 2 : row_bytes(dt, cx));
        s.nblk = dt == DT_F16 ? cx / 32 : cx / block_capacity(dt);
        int target = STEP_TARGET_UNIT;
        int R = pick_rows(segs, nseg, s.row_bytes, nmat, true, sms, target);
        if (R == 0) { target = 2 * STEP_TARGET_UNIT; R = pick_rows(segs, nseg, s.row_bytes, nmat, true, sms, target); }
        IFB_CHECK_ARG(R > 0, "step kernel: rows of op %d (row_bytes %d) cannot be tiled into ring units", kind, s.row_bytes);
        s.R = R;
        s.tasks = R / 2;                         // two adjacent rows (two adjacent w1,w3 row pairs for GLU) per task
        s.unit_bytes = R * s.row_bytes;
        int units = 0;
        s.seg1 = s.seg2 = 0x7fffffff;
        for (int i = 0; i < nseg; i++) {
            if (i == 1) s.seg1 = units;
            if (i == 2) s.seg2 = units;
            units += segs[i] / R;
        }
        s.units = units;
        s.vec4 = (dt == DT_Q4_B32T1A || dt == DT_Q4_B32T1B) && cx % 128 == 0;
        s.nxb = cx / 32;
        if (s.vec4) s.nxb = (s.nxb + 127) / 128 * 128;
        IFB_CHECK_ARG(cx % 32 == 0 && (dt == DT_F16 || cx % cap == 0), "step kernel: cx %d of op %d is not a multiple of the block capacity", cx, kind);
        return IFB_OK;
    };
    int rc;
    { const int segs[3] = {P.q_dim, P.kv_dim, P.kv_dim}; if ((rc = gemv_shape(OPK_QKV, segs, 3, P.d, wdt, 1)) != IFB_OK) return rc; }
    { const int segs[1] = {P.d}; if ((rc = gemv_shape(OPK_WO, segs, 1, P.q_dim, wdt, 1)) != IFB_OK) return rc; }
    { const int segs[1] = {P.ffn}; if ((rc = gemv_shape(OPK_GLU, segs, 1, P.d, wdt, 2)) != IFB_OK) return rc; }
    { const int segs[1] = {P.d}; if ((rc = gemv_shape(OPK_W2, segs, 1, P.ffn, wdt, 1)) != IFB_OK) return rc; }
    { const int segs[1] = {P.vocab_l}; if ((rc = gemv_shape(OPK_HEAD, segs, 1, P.d, head_dtype, 1)) != IFB_OK) return rc; }
    // ---- warp-private pipelines (P.priv requested by the caller): eligible when every chunk fits a private slot
    if (P.priv) {
        const char *why = nullptr;
        auto priv_shape = [&](int kind, const int *segs, int nseg, int dt) {
            const StepOpShape &s = P.op[kind];
            StepPrivShape &v = P.pv[kind];
            v = StepPrivShape();
            const int rb = s.row_bytes;
            int rt = 2;
            while (rt <= 4 && (rt * rb) % 16 != 0) rt *= 2;
            if (rt > 4) { why = "rows are not a multiple of 4 bytes"; return; }
            int total = 0;
            for (int i = 0; i < nseg; i++) { total += segs[i]; if (segs[i] % rt) why = "a matrix has a row count that is no multiple of the task's rows"; }
            v.rt = rt;
            v.ntasks = total / rt;
            if (rt * rb <= STEP_PRIV_CHUNK) {
                v.nc = 1; v.kc_bytes = rb; v.kc_blk = dt == DT_F16 ? s.cx : s.nblk;
                return;
            }
            if (rt != 2 || rb % 16 != 0) { why = "long rows that are no multiple of 16 bytes"; return; }
            const int half = STEP_PRIV_CHUNK / 2;
            if (dt == DT_F16) {
                int nc = (rb + half - 1) / half;
                int ke = ((s.cx + nc - 1) / nc + 7) / 8 * 8;        // halves per chunk: equal chunks, 16-byte granules
                v.nc = (s.cx + ke - 1) / ke; v.kc_blk = ke; v.kc_bytes = ke * 2;
            } else if (s.vec4) {
                v.kc_blk = 128; v.kc_bytes = 128 * block_bytes(dt); v.nc = (s.nblk + 127) / 128;   // one pass of the lane-owns-4-blocks loop
                if (v.kc_bytes > half) why = "a 128-block pass does not fit half a private slot";
            } else {
                const int bb = block_bytes(dt);
                int nc = (rb + half - 1) / half, kb = 0;
                for (;; nc++) {
                    kb = ((s.nblk + nc - 1) / nc + 31) / 32 * 32;   // blocks per chunk: whole lane strides
                    if (kb * bb <= half || kb <= 32) break;
                }
                if (kb * bb > half) { why = "32 blocks do not fit half a private slot"; return; }
                v.kc_blk = kb; v.kc_bytes = kb * bb; v.nc = (s.nblk + kb - 1) / kb;
            }
        };
        { const int segs[3] = {P.q_dim, P.kv_dim, P.kv_dim}; priv_shape(OPK_QKV, segs, 3, wdt); }
        { const int segs[1] = {P.d}; priv_shape(OPK_WO, segs, 1, wdt); }
        { const int segs[1] = {P.ffn}; priv_shape(OPK_GLU, segs, 1, wdt); }
        { const int segs[1] = {P.d}; priv_shape(OPK_W2, segs, 1, wdt); }
        { const int segs[1] = {P.vocab_l}; priv_shape(OPK_HEAD, segs, 1, head_dtype); }
        P.pv[OPK_ATT] = StepPrivShape();
        const int att_chunk = P.kv_q8 ? 32 * (P.hd + (P.hd / 32) * 2) : 32 * P.hd * 2;   // one plane of a 32-position tile
        if (att_chunk > STEP_PRIV_CHUNK) why = "a 32-position cache tile does not fit a private slot";
        if (why == nullptr) {
            P.att_pos = 32;
            P.pslot = (STEP_PRIV_CHUNK + 16 + 127) / 128 * 128;   // + 16: Q8_B32T2 block decode reads one word past its block
        } else {
            P.priv = 0;          // the shared-ring layout below takes it
        }
    }
    StepOpShape &a = P.op[OPK_ATT];
    a = StepOpShape();
    a.tasks = P.att_pos / 32;
    a.unit_bytes = P.kv_q8 ? P.att_pos * (2 * P.hd + 2 * (P.hd / 32) * 2) : P.att_pos * P.hd * 4;

    // shared memory: ring | staged x | residual stream | per-warp score scratch | RoPE table
    int64_t slot = a.unit_bytes;
    for (int k = 0; k < OPK_COUNT; k++) {
        if (k == OPK_ATT) continue;
        const int64_t ub = (int64_t)P.op[k].unit_bytes * (k == OPK_GLU ? 2 : 1);
        if (ub > slot) slot = ub;
    }
    slot = (slot + 16 + 127) / 128 * 128;        // + 16: Q8_B32T2 block decode reads one word past its block
    int64_t xb = att_layout(P.heads, P.hd, P.kv_q8 != 0).total;
    for (int k = 0; k < OPK_COUNT; k++) {
        if (k == OPK_ATT) continue;
        const bool f16x = (k == OPK_HEAD ? head_dtype : wdt) == DT_F16;
        const int64_t b = f16x ? (int64_t)P.op[k].cx * 2 : (int64_t)P.op[k].nxb * 40;
        if (b > xb) xb = b;
    }
    const int64_t merge_b = (int64_t)(STEP_CTHREADS / (P.hd / 4)) * (P.hd + 4) * 4;
    if (merge_b > xb) xb = merge_b;
    xb = (xb + 127) / 128 * 128;
    const int64_t resid_b = ((int64_t)P.d * 2 + 127) / 128 * 128;
    const int64_t sp_b = ((int64_t)STEP_CWARPS * 32 * 4 + 127) / 128 * 128;
    const int64_t rope_b = ((int64_t)P.hd * 4 + 127) / 128 * 128;
    const int64_t lay_b = ((int64_t)P.n_layers * (int64_t)sizeof(StepLayer) + 127) / 128 * 128;
    const int64_t norm_b = resid_b;
    const int64_t smem_max = 232448 - 1536;     // opt-in maximum per CTA minus this kernel's static shared memory
    int64_t nslots = (smem_max - xb - resid_b - sp_b - rope_b - lay_b - norm_b) / slot;
    if (P.priv) {
        if (nslots * slot >= (int64_t)2 * STEP_CWARPS * P.pslot) {
            nslots = 2 * STEP_CWARPS;
            slot = P.pslot;
        } else {
            P.priv = 0;          // the staging areas of this model leave no room for 30 private slots
        }
    }
    if (!P.priv && nslots > STEP_MAX_SLOTS) nslots = STEP_MAX_SLOTS;
    IFB_CHECK_ARG(nslots >= 3, "step kernel: only %lld ring slots of %lld bytes fit beside %lld bytes of staging", (long long)nslots, (long long)slot, (long long)(xb + resid_b));
    P.nslots = (int)nslots;
    P.slot_bytes = (int)slot;
    P.x_off = (int)(nslots * slot);
    P.resid_off = P.x_off + (int)xb;
    P.sp_off = P.resid_off + (int)resid_b;
    P.rope_off = P.sp_off + (int)sp_b;
    P.lay_off = P.rope_off + (int)rope_b;
    P.norm_off = P.lay_off + (int)lay_b;
    P.smem_bytes = P.norm_off + (int)norm_b;
    return IFB_OK;
}

template <class Fmt, bool PRIV>
static int launch_fmt2(const StepParams &P, int grid, cudaStream_t st) {
    auto kern = decode_step_kernel<Fmt, PRIV>;
    static int attr_dev = -1;   // per instantiation: the opt-in is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        IFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1536));
        attr_dev = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(STEP_THREADS);
    cfg.dynamicSmemBytes = (size_t)P.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // every CTA must be resident: they wait for each other
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    IFB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    return IFB_OK;
}

template <class Fmt>
static int launch_fmt(const StepParams &P, int grid, cudaStream_t st) {
    return P.priv ? launch_fmt2<Fmt, true>(P, grid, st) : launch_fmt2<Fmt, false>(P, grid, st);
}

int step_launch(const StepParams &P, int grid, cudaStream_t st) {
    switch (P.w_dtype) {
        case DT_F16: return launch_fmt<FmtF16>(P, grid, st);
        case DT_Q4_B32T1A:
        case DT_Q4_B32T1B: return launch_fmt<FmtQ4_B32T1>(P, grid, st);
        case DT_Q4_B64T1: return launch_fmt<FmtQ4_B64T1>(P, grid, st);
        case DT_Q8_B32T2: return launch_fmt<FmtQ8_B32T2>(P, grid, st);
        case DT_Q6_B64T1: return launch_fmt<FmtQ6_B64T1>(P, grid, st);
        case DT_Q5_B64T1: return launch_fmt<FmtQ5_B64T1>(P, grid, st);
        case DT_Q3H_B64T1: return launch_fmt<FmtQ3H_B64T1>(P, grid, st);
        default:
            set_last_error("step kernel: weight type %d", P.w_dtype);
            return IFB_ERR_TYPE;
    }
}

}  // namespace ifb
