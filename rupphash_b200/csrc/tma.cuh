// tma.cuh -- the sm_100a bulk-copy engine (TMA) and mbarrier primitives the kernels use, as
// inline PTX: 1-D cp.async.bulk global -> shared with mbarrier transaction-byte completion
// (SASS: UBLKCP + SYNCS), and the debug-build index assertion.
#pragma once
#include <stdint.h>

namespace rh {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// makes the initialised barriers visible to the async proxy (the copy engine) before first use
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one arrival + `bytes` of expected copy-engine traffic
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bulk copy global -> shared; `bytes` is a multiple of 16, both addresses 16-byte aligned.  The
// copy engine signals `bar` with the byte count when the data has landed.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic-proxy writes (global and shared) of this thread -> ordered before later async-proxy accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// generic-proxy writes to shared memory -> visible to / ordered before later async-proxy accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace rh

// RH_DEBUG_BOUNDS build (make debug): every shared / global index the hot kernels compute is checked
// in the kernel and a violation traps (the launch then fails with an error the ABI reports) -- the
// substitute for compute-sanitizer, which is closed on the GPU pool.
#ifdef RH_DEBUG_BOUNDS
#define RH_CHECK_IDX(i, n)                                                                   \
    do {                                                                                     \
        if (!((long long)(i) >= 0 && (long long)(i) < (long long)(n))) {                     \
            printf("RH_DEBUG_BOUNDS %s:%d: index %lld outside [0, %lld)\n", __FILE__, __LINE__, \
                   (long long)(i), (long long)(n));                                          \
            __trap();                                                                        \
        }                                                                                    \
    } while (0)
#else
#define RH_CHECK_IDX(i, n) \
    do {                   \
    } while (0)
#endif
