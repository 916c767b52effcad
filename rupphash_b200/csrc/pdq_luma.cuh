// pdq_luma.cuh -- the packed-pixel luma front end shared by the fused PDQ kernel (pdq_fused.cu) and
// the generic pipeline's vectorised luma kernel (pdq.cu): 128-bit chunk loads, luma601 with two DP2A
// per pixel and the division by 1000 as one FP32 FMA, the 2x Box pre-downsample as two rounded
// halving passes (pdqhash.rs:203-220, :268-284).
#pragma once
#include <stdint.h>

#include "../../include/rupphash_b200.h"

namespace rh {

// L2 eviction policies (createpolicy): pixels are read exactly once, so they are marked evict-first and
// leave the L2 to the data that is re-read (the fused kernel's pass-3 slab).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 128-bit read-only load with an L2 cache hint
__device__ __forceinline__ uint4 ldg_hint(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

template <int BYTES>
__device__ __forceinline__ void load_chunk_hint(const uint8_t *p, uint32_t *w, uint64_t pol) {
    static_assert(BYTES % 16 == 0, "hinted loads are 128-bit");
#pragma unroll
    for (int i = 0; i < BYTES / 16; i++) {
        uint4 v = ldg_hint(reinterpret_cast<const uint4 *>(p) + i, pol);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
}

template <int BYTES>
__device__ __forceinline__ void load_chunk(const uint8_t *p, uint32_t *w) {
    if (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; i++) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    } else if (BYTES % 8 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 8; i++) {
            uint2 v = __ldg(reinterpret_cast<const uint2 *>(p) + i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < BYTES / 4; i++) w[i] = __ldg(reinterpret_cast<const uint32_t *>(p) + i);
    }
}

// pdqhash.rs:268-284 for pixel k of a chunk held in words w[].  RGB pixels straddle words; every
// alignment is two DP2A (16-bit weights x 8-bit samples) with no byte shuffling.
//
// The division by 1000 runs on the FP32 pipe instead of the (two-pass) IMAD.HI: the DP2A pair
// accumulates 2v + 1001 on top of LUMA_F0, the bit pattern of the float 8390000 = 2^23 + 1392, so
// its result IS the float F = 8390000 + (2v + 1001) (exact: < 2^24).  8390000 / 2000 = 4195, so
// RN(F * 0.0005f + (2^23 - 4195)) = 2^23 + RN((2v + 1001) / 2000) = 2^23 + floor(v / 1000) + 1:
// 2v + 1001 is odd, hence never closer than 1/2000 to a rounding tie, while the error of the
// product is < 2.2e-4 (exhaustively checked for every v <= 256000, tests/test_fused_model.py).
// The return value is LUMA_K + luma + 1; callers fold the bias into their next integer add.
constexpr uint32_t LUMA_K = 0x4B000000u;                         // bits of 2^23
constexpr uint32_t LUMA_F0 = LUMA_K + 1392u + 1001u + 2u * 500u;   // + 2 * (the reference's +500 rounding term)
template <int LAYOUT>
__device__ __forceinline__ uint32_t luma_px(const uint32_t *w, int k) {
    constexpr uint32_t W_RG = 598u | (1174u << 16), W_B0 = 228u, W_0R = 598u << 16, W_GB = 1174u | (228u << 16);
    if (LAYOUT == RH_LAYOUT_LUMA8) return LUMA_K + 1u + ((w[k >> 2] >> (8 * (k & 3))) & 0xFFu);
    uint32_t v;
    if (LAYOUT == RH_LAYOUT_RGBA8) {
        v = __dp2a_lo(W_RG, w[k], LUMA_F0);
        v = __dp2a_hi(W_B0, w[k], v);
    } else {
        const int o = 3 * k, i = o >> 2, sh = o & 3;
        if (sh == 0) {
            v = __dp2a_lo(W_RG, w[i], LUMA_F0);
            v = __dp2a_hi(W_B0, w[i], v);
        } else if (sh == 1) {
            v = __dp2a_lo(W_0R, w[i], LUMA_F0);
            v = __dp2a_hi(W_GB, w[i], v);
        } else if (sh == 2) {
            v = __dp2a_hi(W_RG, w[i], LUMA_F0);
            v = __dp2a_lo(W_B0, w[i + 1], v);
        } else {
            v = __dp2a_hi(W_0R, w[i], LUMA_F0);
            v = __dp2a_lo(W_GB, w[i + 1], v);
        }
    }
    return __float_as_uint(__fmaf_rn(__uint_as_float(v), 0.0005f, 8384413.0f));
}

// 8 consecutive luma pixels of one plane row from the thread's source chunk(s).
template <int LAYOUT, bool DOWN2, int NW>
__device__ __forceinline__ uint2 luma8(const uint32_t *w0, const uint32_t *w1) {
    uint32_t l[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (DOWN2) {
            // Box 2x: horizontal pass first, each pass rounds half up (pdqhash.rs:203-220)
            constexpr uint32_t BIAS2 = 0u - (2u * LUMA_K + 1u);   // two biased pixels -> a + b + 1
            const uint32_t h0 = (luma_px<LAYOUT>(w0, 2 * k) + luma_px<LAYOUT>(w0, 2 * k + 1) + BIAS2) >> 1;
            const uint32_t h1 = (luma_px<LAYOUT>(w1, 2 * k) + luma_px<LAYOUT>(w1, 2 * k + 1) + BIAS2) >> 1;
            l[k] = (h0 + h1 + 1u) >> 1;
        } else {
            l[k] = luma_px<LAYOUT>(w0, k) - (LUMA_K + 1u);
        }
    }
    uint2 r;
    r.x = l[0] | (l[1] << 8) | (l[2] << 16) | (l[3] << 24);
    r.y = l[4] | (l[5] << 8) | (l[6] << 16) | (l[7] << 24);
    return r;
}

}  // namespace rh
