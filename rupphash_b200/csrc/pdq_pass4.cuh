// pdq_pass4.cuh -- pieces shared by the two fused PDQ kernels (pdq_fused.cu: planes 512 wide, integer
// passes 1-2; pdq_float.cu: every other plane up to 512 x 512, float passes): policy-carrying pixel loads,
// the pass-3 sample slab, pass 4 (column chains over the slab + decimation) and the phase clock.
// Both kernels run 256 threads per CTA.
#pragma once
#include "pdq_luma.cuh"
#include "pdq_tail.cuh"
#include "tma.cuh"

namespace rh {

constexpr int FTHREADS = 256;     // 8 warps: front end, chains, tail
constexpr int P3_PITCH = 512;     // floats per column of the pass-3 slab ([64][P3_PITCH], column-major)

enum { PH_FRONT = 0, PH_EDGE, PH_CHAIN, PH_P4_STAGE, PH_P4_CHAIN, PH_TAIL, NPHASE };

// pixels are read exactly once: 128-bit loads carry the caller's L2 policy (evict-first by default)
template <int BYTES>
__device__ __forceinline__ void load_px(const uint8_t *p, uint32_t *w, uint64_t pol) {
    if (BYTES % 16 == 0)
        load_chunk_hint<BYTES % 16 == 0 ? BYTES : 16>(p, w, pol);
    else
        load_chunk<BYTES>(p, w);
}

// pass-3 samples go to the per-CTA slab in L2 and come back for pass 4: kept with evict-last priority
__device__ __forceinline__ void st_slab(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// ------------------------------------------------------------------------ tail ----

// Pass 4: the column chains over the pass-3 samples (window WC, length H) for the 64 decimated
// columns, keeping the 64 decimated rows (pdqhash.rs:435).  The slab is pulled from L2 into shared
// memory P4_ROWS rows at a time by the bulk-copy engine (one 512-byte cp.async.bulk per column, mbarrier
// completion) into two buffers: chunk c + 1 lands while threads 0..63 (one per column) walk chunk c at
// shared-memory latency, so only the first chunk's L2 round trip is exposed.
// The pitch is a multiple of 4 floats with pitch / 4 odd: the 128-bit accesses of the walk (8 lanes
// per wavefront) are conflict-free.
// rows in the clipped window of output o of a line of `len` samples (pdqhash.rs:375, :383, :392)
__device__ __forceinline__ int window_count(int o, int len, int ht, int hb) {
    return min(o + hb, len - 1) - max(o - ht, 0) + 1;
}

// sum / count as the reference computes it (an IEEE f32 division); exact scaling for powers of two
__device__ __forceinline__ float div_count(float sum, int cnt) {
    if ((cnt & (cnt - 1)) == 0) return __fmul_rn(sum, 1.0f / (float)cnt);
    return __fdiv_rn(sum, (float)cnt);
}

// How the slab reaches pass 4: already divided (pdq_fused.cu), or as the RAW pass-3 running sums of the
// decimated columns of a plane W wide with row window wr (pdq_float.cu), which pass 4 divides by their window
// sizes as each staged chunk lands -- in shared memory, by all 256 threads (thread = a quarter of a column).
struct SlabNorm {
    int W, wr;   // W == 0: the slab holds quotients already
};

constexpr int P4_ROWS = 128;
constexpr int P4_PITCH = P4_ROWS + 4;
static_assert((P4_PITCH / 4) % 2 == 1 && P4_PITCH % 4 == 0, "pass-4 staging pitch");
static_assert((64 * (P4_ROWS / 4)) % FTHREADS == 0 && P4_ROWS % 8 == 0, "pass-4 staging has no remainder");
constexpr size_t P4_STAGE_OFF = (sizeof(TailSmem) + 15) & ~size_t(15);   // 16-byte aligned for the 128-bit accesses
constexpr size_t P4_SMEM_BYTES = P4_STAGE_OFF + 2 * 64 * P4_PITCH * 4;   // tail scratch + the two staging buffers

struct PhaseClock {
    unsigned long long *acc;
    long long t;
    __device__ __forceinline__ void start(unsigned long long *p) {
        acc = p;
        if (acc != nullptr && threadIdx.x == 0) t = clock64();
    }
    // call right after the barrier that ends phase `ph`
    __device__ __forceinline__ void lap(int ph) {
        if (acc != nullptr && threadIdx.x == 0) {
            const long long now = clock64();
            atomicAdd(acc + ph, (unsigned long long)(now - t));
            t = now;
        }
    }
};

// Running window sums of pass 4 for one staged chunk, in place: on return colp[i] holds the window
// sum after plane row c0 + i has entered (= the sum of output row c0 + i - HB).  Nothing but the
// dependent add / subtract pair per row sits on the chain: decimation and the division by the row
// count are done afterwards by the whole CTA (p4_gather).
template <int WC>
__device__ __forceinline__ void p4_walk(float *colp, int c0, int rows, float &sum, float (&prev)[8]) {
    float cur[8];
    int r0 = 0;
    if (c0 == 0) {   // rows 0 .. 7: the window is still filling for ri < WC (pdqhash.rs:366-378)
        const float4 lo = *reinterpret_cast<const float4 *>(colp);
        const float4 hi = *reinterpret_cast<const float4 *>(colp + 4);
        cur[0] = lo.x; cur[1] = lo.y; cur[2] = lo.z; cur[3] = lo.w;
        cur[4] = hi.x; cur[5] = hi.y; cur[6] = hi.z; cur[7] = hi.w;
        float sums[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            sum = __fadd_rn(sum, cur[k]);
            if (k >= WC) sum = __fsub_rn(sum, cur[k - WC]);
            sums[k] = sum;
        }
        *reinterpret_cast<float4 *>(colp) = make_float4(sums[0], sums[1], sums[2], sums[3]);
        *reinterpret_cast<float4 *>(colp + 4) = make_float4(sums[4], sums[5], sums[6], sums[7]);
#pragma unroll
        for (int k = 0; k < 8; k++) prev[k] = cur[k];
        r0 = 8;
    }
    // steady state (pdqhash.rs:380-387); rows past the image in the last batch are computed on
    // whatever the staging left there and never read
    // (the next batch is loaded before the current one is summed: no LDS latency between batches; the
    // read one batch past the chunk stays inside the staging area)
    float4 lo = *reinterpret_cast<const float4 *>(colp + r0);
    float4 hi = *reinterpret_cast<const float4 *>(colp + r0 + 4);
#pragma unroll 1
    for (; r0 < rows; r0 += 8) {
        cur[0] = lo.x; cur[1] = lo.y; cur[2] = lo.z; cur[3] = lo.w;
        cur[4] = hi.x; cur[5] = hi.y; cur[6] = hi.z; cur[7] = hi.w;
        lo = *reinterpret_cast<const float4 *>(colp + r0 + 8);
        hi = *reinterpret_cast<const float4 *>(colp + r0 + 12);
        float sums[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float old = (k >= WC) ? cur[k - WC] : prev[8 + k - WC];
            sum = __fsub_rn(__fadd_rn(sum, cur[k]), old);
            sums[k] = sum;
        }
        *reinterpret_cast<float4 *>(colp + r0) = make_float4(sums[0], sums[1], sums[2], sums[3]);
        *reinterpret_cast<float4 *>(colp + r0 + 4) = make_float4(sums[4], sums[5], sums[6], sums[7]);
#pragma unroll
        for (int k = 0; k < 8; k++) prev[k] = cur[k];
    }
}

// Decimated rows whose window sum lies in the staged chunk (or in the shrink-phase sums `shr`, for the
// last HB output rows) -> B, divided by the number of rows in the clipped window (pdqhash.rs:372,
// :385, :393: the reference divides the running sum by its running count).
template <int WC>
__device__ __forceinline__ void p4_gather(const float *stage, const float *shr, int H, int c0, int rows, bool last, float *B) {
    constexpr int HALF = (WC + 2) / 2, HB = HALF - 1, HT = WC - HALF;
    const int j = threadIdx.x & 63;
    // first output whose window sum can lie in this chunk: o + HB >= c0  <=  (2 i + 1) H >= 128 (c0 - HB)
    const int i_lo = max(0, (128 * (c0 - HB) - H) / (2 * H));
    for (int i = i_lo + (threadIdx.x >> 6); i < 64; i += FTHREADS / 64) {
        const int o = ((2 * i + 1) * H) >> 7;   // decimated output row (pdqhash.rs:435)
        if (!last && o + HB - c0 >= rows) break;   // the rest belongs to later chunks
        const int src = o + HB - c0;            // staged slot whose sum is output row o
        const float cnt = (float)(min(o + HB, H - 1) - max(o - HT, 0) + 1);
        if (o >= H - HB) {
            if (last) B[i * 64 + j] = __fdiv_rn(shr[(o - (H - HB)) * 64 + j], cnt);
        } else if (src >= 0 && src < rows) {
            B[i * 64 + j] = __fdiv_rn(stage[j * P4_PITCH + src], cnt);
        }
    }
}

// plane rows [c0, c0 + P4_ROWS) of the 64 columns -> stg (rows past H are copied but never used;
// the clamp keeps the last chunk inside its column)
// plane rows [c0, c0 + P4_ROWS) of the 64 slab columns -> stg, by the bulk-copy engine: one 512-byte
// cp.async.bulk per column (threads 0..63, each arming the stage's mbarrier with its own byte count), instead of
// 2048 16-byte cp.async requests per chunk.  c0 is a multiple of P4_ROWS and P3_PITCH = 4 P4_ROWS, so a chunk never
// leaves its column (rows past H are copied but never used).
static_assert(P3_PITCH % P4_ROWS == 0, "a staged chunk stays inside its slab column");
__device__ __forceinline__ void p4_issue(const float *p3t, int c0, float *stg, uint64_t *bar) {
    if (threadIdx.x < 64) {
        const int col = threadIdx.x;
        RH_CHECK_IDX(c0 + P4_ROWS - 1, P3_PITCH);
        mbar_arrive_expect_tx(bar, P4_ROWS * 4);
        bulk_g2s(stg + col * P4_PITCH, p3t + (size_t)col * P3_PITCH + c0, P4_ROWS * 4, bar);
    }
}

template <int WC>
__device__ __forceinline__ void pass4(const float *p3t, int H, float *B, float *stage, float *shr, unsigned long long *bars,
                                      PhaseClock &clk, const SlabNorm norm = SlabNorm{0, 0}) {
    constexpr int HALF = (WC + 2) / 2, HB = HALF - 1;
    const int j = threadIdx.x;
    float sum = 0.0f;
    float prev[8];
#pragma unroll
    for (int k = 0; k < 8; k++) prev[k] = 0.0f;
    uint64_t *bar = reinterpret_cast<uint64_t *>(bars);
    // raw slab: this thread's column (thread = column j for the walk; column tid / 4 for the chunk pass) and divisor
    int cnt_walk = 1, cnt_chunk = 1;
    if (norm.W) {
        const int half = (norm.wr + 2) / 2, ht = norm.wr - half, hb = half - 1;
        cnt_walk = window_count(((2 * (j & 63) + 1) * norm.W) >> 7, norm.W, ht, hb);
        cnt_chunk = window_count(((2 * (j >> 2) + 1) * norm.W) >> 7, norm.W, ht, hb);
    }
    // the slab was written with ordinary stores by this CTA: order them before the copy engine's reads, and set up
    // the two stage barriers (64 arrivals: one per issuing thread) in memory the band buffers used until now
    fence_proxy_async_all();
    if (j == 0) {
        mbar_init(&bar[0], 64);
        mbar_init(&bar[1], 64);
        mbar_fence_init();
    }
    __syncthreads();
    p4_issue(p3t, 0, stage, &bar[0]);
    if (P4_ROWS < H) p4_issue(p3t, P4_ROWS, stage + 64 * P4_PITCH, &bar[1]);
    int buf = 0;
    uint32_t parity = 0;   // bit b: phase of stage b's barrier
    for (int c0 = 0; c0 < H; c0 += P4_ROWS, buf ^= 1) {
        float *stg = stage + buf * (64 * P4_PITCH);
        const int rows = min(P4_ROWS, H - c0);
        const bool last = c0 + P4_ROWS >= H;
        // the rows that leave during the shrink phase (H-WC .. H-WC+HB-1), fetched under the walk
        float leave[HB > 0 ? HB : 1];
        if (last && j < 64) {
#pragma unroll
            for (int k = 0; k < HB; k++) {
                leave[k] = __ldcg(p3t + (size_t)j * P3_PITCH + (H - WC + k));
                if (norm.W) leave[k] = div_count(leave[k], cnt_walk);
            }
        }
        mbar_wait(&bar[buf], (parity >> buf) & 1u);   // this chunk has landed (the next one is still in flight)
        parity ^= 1u << buf;
        if (norm.W) {   // raw sums -> quotients (pdqhash.rs:375, :383, :392), 32 consecutive rows of one column per thread
            float *seg = stg + (j >> 2) * P4_PITCH + (j & 3) * (P4_ROWS / 4);
            const bool pow2 = (cnt_chunk & (cnt_chunk - 1)) == 0;
            const float fc = (float)cnt_chunk, inv = 1.0f / fc;
#pragma unroll
            for (int k = 0; k < P4_ROWS / 4; k += 4) {
                float4 v = *reinterpret_cast<float4 *>(seg + k);
                v.x = pow2 ? __fmul_rn(v.x, inv) : __fdiv_rn(v.x, fc);
                v.y = pow2 ? __fmul_rn(v.y, inv) : __fdiv_rn(v.y, fc);
                v.z = pow2 ? __fmul_rn(v.z, inv) : __fdiv_rn(v.z, fc);
                v.w = pow2 ? __fmul_rn(v.w, inv) : __fdiv_rn(v.w, fc);
                *reinterpret_cast<float4 *>(seg + k) = v;
            }
            __syncthreads();
        }
        clk.lap(PH_P4_STAGE);
        if (j < 64) {
            p4_walk<WC>(stg + j * P4_PITCH, c0, rows, sum, prev);
            if (last) {   // shrink phase (pdqhash.rs:389-395): outputs H-HB .. H-1
                // the walk may have run past row H-1 inside its last batch of 8: restart from the sum of row H-1
                sum = stg[j * P4_PITCH + rows - 1];
#pragma unroll
                for (int k = 0; k < HB; k++) {
                    sum = __fsub_rn(sum, leave[k]);
                    shr[k * 64 + j] = sum;
                }
            }
            fence_proxy_async();   // the in-place sums are ordered before a later bulk copy into this stage
        }
        __syncthreads();
        clk.lap(PH_P4_CHAIN);
        p4_gather<WC>(stg, shr, H, c0, rows, last, B);
        if (c0 + 2 * P4_ROWS < H) {   // this buffer takes the chunk after the next one
            __syncthreads();
            p4_issue(p3t, c0 + 2 * P4_ROWS, stg, &bar[buf]);
        }
    }
    __syncthreads();
    if (j == 0) {   // the barriers' memory goes back to the band buffers
        mbar_inval(&bar[0]);
        mbar_inval(&bar[1]);
    }
}

}  // namespace rh
