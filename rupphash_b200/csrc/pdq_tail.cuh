// pdq_tail.cuh -- the 64x64 -> hash tail of PDQ shared by every hashing kernel:
//   quality metric      pdqhash.rs:445-460
//   dct64_to_16         pdqhash.rs:306-336   (C = D * B * D^T, k-ascending, mul then add, no FMA)
//   rank-127 median     pdqhash.rs:116-124   (f32::total_cmp order)
//   bits + packing      pdqhash.rs:91-106, :155-162
//   8 dihedral variants pdqhash.rs:71-87, :127-151
// One CTA of 256 threads per image; thread n owns coefficient n = 16 r + c.
#pragma once
#include <stdint.h>

namespace rh {

constexpr int TAIL_THREADS = 256;

// The tail synchronises on named barrier 1 over exactly TAIL_THREADS threads, so that a kernel
// with extra warps can run it on its first 256 threads only.
__device__ __forceinline__ void tail_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ int tail_count(bool pred) {
    int n;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.popc.u32 %0, 1, 256, p;\n\t}"
                 : "=r"(n)
                 : "r"((unsigned)pred)
                 : "memory");
    return n;
}
constexpr int DCT_PITCH = 65;  // padded row pitch of the 16 x 64 DCT matrix in shared memory

struct TailSmem {
    float B[64 * 64];            // decimated buffer, row-major
    float D[16 * DCT_PITCH];     // DCT matrix rows (frequencies 1..16)
    float T[16 * 64];            // D * B
    float C[256];                // coefficients
    uint8_t bits[4][256];        // one bit per coefficient and sign pattern
    int red[8];
    float median;
    unsigned long long p4_bar[2];   // mbarriers of the pass-4 staging buffers (fused kernels only)
};

struct TailOut {
    uint8_t *hash;       // n x 32 or nullptr
    float *quality;      // n or nullptr
    float *coeffs;       // n x 256 or nullptr
    uint8_t *dihedral;   // n x 8 x 32 or nullptr
};

// f32::total_cmp as an unsigned key (ascending)
__device__ __forceinline__ uint32_t total_ukey(float f) {
    int32_t b = __float_as_int(f);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return (uint32_t)b ^ 0x80000000u;
}
__device__ __forceinline__ float from_total_ukey(uint32_t u) {
    int32_t b = (int32_t)(u ^ 0x80000000u);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return __int_as_float(b);
}

// pdqhash.rs:127-137: negate odd FREQUENCIES r+1 / c+1, i.e. even array indices
__device__ __forceinline__ float apply_sign(float v, int r, int c, bool neg_rows, bool neg_cols) {
    bool flip = (neg_rows && ((r & 1) == 0)) != (neg_cols && ((c & 1) == 0));
    return flip ? -v : v;
}

// Element of rank 127 (0-based) of the 256 values held one per thread, under total_cmp.
// MSB-first radix select: 32 block-wide counts.  Every thread returns the same value.
__device__ __forceinline__ float block_rank127(float v) {
    const uint32_t key = total_ukey(v);
    uint32_t prefix = 0, mask = 0;
    int k = 127;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t b = 1u << bit;
        const int c = tail_count(((key & mask) == prefix) && !(key & b));
        if (k >= c) {
            k -= c;
            prefix |= b;
        }
        mask |= b;
    }
    return from_total_ukey(prefix);
}

// write one packed hash: thread n contributes `bit` for coefficient n (pdqhash.rs:155-162:
// coefficient n lands in bit n%8 of byte 31 - n/8)
__device__ __forceinline__ void pack_hash(bool bit, uint8_t *hash) {
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane < 4) hash[31 - 4 * w - lane] = (uint8_t)(ballot >> (8 * lane));
}

// Coefficients in s.C -> hash / dihedral hashes.  All 256 threads must call.
__device__ __forceinline__ void tail_hashes(TailSmem &s, const TailOut &o, size_t img) {
    const int n = threadIdx.x, r = n >> 4, c = n & 15;
    const float v = s.C[n];
    if (!o.dihedral) {
        if (o.hash) {
            const float med = block_rank127(v);
            pack_hash(v > med, o.hash + img * 32);
        }
        return;
    }
    // sign patterns: 0 = id, 1 = neg_cols, 2 = neg_rows, 3 = neg_both (pdqhash.rs:72-75)
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
        const float sv = apply_sign(v, r, c, (p & 2) != 0, (p & 1) != 0);
        const float med = block_rank127(sv);
        s.bits[p][n] = sv > med;
    }
    tail_sync();
    const int nt = 16 * c + r;  // transposed source: bit (r, c) of T(x) is bit (c, r) of x
    uint8_t *out = o.dihedral + img * 256;
    // order of pdqhash.rs:77-86
    pack_hash(s.bits[0][n], out + 0 * 32);    // identity
    pack_hash(s.bits[2][nt], out + 1 * 32);   // T(neg_rows)
    pack_hash(s.bits[3][n], out + 2 * 32);    // neg_both
    pack_hash(s.bits[1][nt], out + 3 * 32);   // T(neg_cols)
    pack_hash(s.bits[1][n], out + 4 * 32);    // neg_cols
    pack_hash(s.bits[2][n], out + 5 * 32);    // neg_rows
    pack_hash(s.bits[0][nt], out + 6 * 32);   // T(id)
    pack_hash(s.bits[3][nt], out + 7 * 32);   // T(neg_both)
    if (o.hash) pack_hash(s.bits[0][n], o.hash + img * 32);
}

// pdqhash.rs:445-460.  Every term is an integer <= 100 and the total <= 806 400 < 2^24, so the
// reference's f32 running sum is exact and order-independent: an integer block reduction gives
// the identical value.  Needs s.B; all threads must call; result valid in thread 0.
__device__ __forceinline__ float tail_quality(TailSmem &s) {
    int acc = 0;
    // (rolled: measured, the fused kernel is faster with this loop small than with it unrolled)
#pragma unroll 1
    for (int idx = threadIdx.x; idx < 4096; idx += TAIL_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        const float a = s.B[idx];
        if (i < 63) {
            float d = __fdiv_rn(__fmul_rn(__fsub_rn(a, s.B[idx + 64]), 100.0f), 255.0f);
            acc += (int)fabsf(d);
        }
        if (j < 63) {
            float d = __fdiv_rn(__fmul_rn(__fsub_rn(a, s.B[idx + 1]), 100.0f), 255.0f);
            acc += (int)fabsf(d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = acc;
    tail_sync();
    float q = 0.0f;
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < TAIL_THREADS / 32; w++) tot += s.red[w];
        q = __fdiv_rn((float)tot, 90.0f);
        if (q > 1.0f) q = 1.0f;
    }
    return q;
}

// pdqhash.rs:306-336 with s.B loaded and the DCT matrix (16 rows of pitch DCT_PITCH, shared memory)
// at D; leaves the coefficients in s.C.
__device__ __forceinline__ void tail_dct(TailSmem &s, const float *D) {
    {   // T[i][j] = sum_k D[i][k] * B[k][j], k ascending from 0.0
        const int j = threadIdx.x & 63, i0 = (threadIdx.x >> 6) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < 64; k++) {
            const float b = s.B[k * 64 + j];
#pragma unroll
            for (int u = 0; u < 4; u++) acc[u] = __fadd_rn(acc[u], __fmul_rn(D[(i0 + u) * DCT_PITCH + k], b));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) s.T[(i0 + u) * 64 + j] = acc[u];
    }
    tail_sync();
    {   // C[i][j] = sum_k T[i][k] * D[j][k]
        const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; k++) acc = __fadd_rn(acc, __fmul_rn(s.T[i * 64 + k], D[j * DCT_PITCH + k]));
        s.C[threadIdx.x] = acc;
    }
    tail_sync();
}

}  // namespace rh
