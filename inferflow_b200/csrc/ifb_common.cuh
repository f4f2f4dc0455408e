// inferflow_b200 -- common device/host helpers for the sm_100a kernels.
// PTX wrappers: mbarrier, cp.async.bulk (TMA 1-D bulk copies; SASS UBLKCP),
// programmatic dependent launch (griddepcontrol), warp reductions.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define IFB_WARP 32

// ---------------------------------------------------------------------------
// error handling (host side). Every C-ABI entry returns 0 or a negative code
// and leaves a message retrievable with ifb_last_error().
// ---------------------------------------------------------------------------
namespace ifb {
void set_last_error(const char *fmt, ...);
}

#define IFB_OK 0
#define IFB_ERR_ARG (-1)       // bad argument / unsupported shape (reference: LogError + return false)
#define IFB_ERR_TYPE (-2)      // unsupported element type
#define IFB_ERR_CUDA (-3)      // CUDA runtime error
#define IFB_ERR_STATE (-4)     // object used in the wrong state
#define IFB_ERR_IO (-5)        // file / config error

#define IFB_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            ifb::set_last_error(__VA_ARGS__);         \
            return IFB_ERR_ARG;                       \
        }                                             \
    } while (0)

#define IFB_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            ifb::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,      \
                                cudaGetErrorString(e_));                           \
            return IFB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

#define IFB_LAUNCH_CHECK()                                                         \
    do {                                                                           \
        cudaError_t e_ = cudaGetLastError();                                       \
        if (e_ != cudaSuccess) {                                                   \
            ifb::set_last_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,  \
                                cudaGetErrorString(e_));                           \
            return IFB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

namespace ifb {

// ---------------------------------------------------------------------------
// device: shared-memory address / mbarrier / bulk copy
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    // make the inits visible to the async proxy (TMA) before the first copy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// the same with a suspend-time hint (ns): the thread sleeps in the barrier unit until the phase completes or the hint
// expires, instead of returning at once and spinning through issue slots
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}

// variants on a pre-computed shared-window address (the generic->shared conversion is not free in a hot loop)
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint_a(uint32_t dst_addr, const void *src_gmem, uint32_t bytes, uint32_t bar_addr,
                                                uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst_addr),
        "l"(src_gmem), "r"(bytes), "r"(bar_addr), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D TMA bulk copy global -> shared, completion on an mbarrier (complete_tx).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// same with an L2 evict-first policy (weights / KV are streamed once per token)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem,
                                              uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// L2 prefetch of a contiguous range (no smem destination)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes)
                 : "memory");
}

// ---------------------------------------------------------------------------
// programmatic dependent launch (PDL)
// ---------------------------------------------------------------------------
// Wait until the grids this one depends on have completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Allow dependent grids to start launching (their pre-wait prologue only).
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// optional timeline probe: trace[cta*8 + slot] = ns timestamp (first 160 CTAs of a launch)
__device__ __forceinline__ void trace_mark(unsigned long long *trace, int slot) {
    if (trace != nullptr && blockIdx.x < 160 && blockIdx.y == 0) trace[blockIdx.x * 8 + slot] = globaltimer_ns();
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float h2f(uint16_t bits) {
    return __half2float(__ushort_as_half(bits));
}
__device__ __forceinline__ uint16_t f2h_bits(float v) {
    return __half_as_ushort(__float2half_rn(v));
}
// round-trip through fp16 (the reference stores every intermediate tensor as half)
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// system-scope flag traffic between GPUs (peer memory over NVLink)
__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 128-bit load that never hits a (possibly stale) L1 line: data a peer GPU wrote into this GPU's memory
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void *p, uint4 v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 ld_relaxed_sys_v2(const void *p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Every wait on memory another GPU writes is BOUNDED: a peer that died (exception between tp_connect and the last
// barrier, out of memory, a skipped step) must not wedge the other GPUs of the box in an unkillable kernel. Past the
// deadline the kernel traps; the host sees a launch failure at its next synchronisation and can tear down.
#ifndef IFB_PEER_DEADLINE_NS
#define IFB_PEER_DEADLINE_NS 2000000000ull
#endif
struct PeerDeadline {
    unsigned long long t0;
    unsigned it;
    __device__ __forceinline__ PeerDeadline() : t0(0ull), it(0u) {}
    // call once per failed poll
    __device__ __forceinline__ void tick() {
        if ((++it & 1023u) != 0u) return;
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0ull) t0 = now;
        else if (now - t0 > IFB_PEER_DEADLINE_NS) __trap();
    }
};

// q = (int)roundf(v / scale) exactly as the reference's alg-2 quantizer computes it
// (kernels/tensor_quant.h:64), but cheaply: r = v * (1/scale) is within ~3e-5 of the correctly rounded
// quotient for |q| <= 127; round it to nearest with the 1.5*2^23 trick and take the IEEE division +
// roundf path only when r sits within 2e-4 of a .5 tie (where the two roundings could disagree).
__device__ __forceinline__ int q8_round_div(float v, float scale, float inv_scale) {
    const float r = v * inv_scale;
    const float y = r + 12582912.0f;              // 1.5 * 2^23: the low mantissa bits now hold rne(r)
    const float fl = y - 12582912.0f;
    int q = __float_as_int(y) - 0x4b400000;
    if (fabsf(r - fl) > 0.4998f) q = (int)roundf(__fdiv_rn(v, scale));
    return q;
}

// mixed-sign dp4a: a holds 4 unsigned bytes (weight codes 0..255), b 4 signed bytes
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

}  // namespace ifb
