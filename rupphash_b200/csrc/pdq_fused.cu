// pdq_fused.cu -- the bandwidth kernel of hot path #1: RGB8/RGBA8/Luma8 pixels -> luma601 ->
// (2x Box pre-downsample) -> two Jarosz box-filter repetitions -> 64x64 decimation -> quality,
// DCT, median, hash, in ONE persistent kernel for planes 512 px wide and 193..512 px high (both
// BASELINE shapes: 1024x768 -> 512x384 and 512x512; 16:9 -> 512x288 ...).  Replaces
// pdqhash.rs:166-262 for those shapes; every other shape takes the generic pipeline in pdq.cu.
// Results are bit-identical to the reference's sequential float arithmetic (tools/fused_model.py
// proves the restructuring on the CPU).
//
// Why it is not four float passes.  With a row window of 8 (pdqhash.rs:246: ceil(512/64)):
//   * pass 1 (rows) of u8 luma is exact -- P1 = H/8, H an integer <= 2040 -- except in the six
//     columns 0,1,2,508,509,510 whose clipped windows divide by 5,6,7;
//   * pass 2 (columns) of exact eighths has exact running sums, hence away from those columns
//     P2[r][c] = RN(S2d / (8 cnt_r)) with S2d the integer 2-D box sum: integer work + 1 rounding;
//   * passes 3 and 4 are true sequential float chains (rounding error persists along a line) and
//     are run as written -- one thread per row / per column -- but pass 4 only needs the 64
//     decimated columns of pass 3, so only those are kept (64 floats per row).
// The six inexact columns get the reference's real column chain (one lane each).
//
// Data flow.  pdq_edge_kernel first runs the six inexact columns of every image of the chunk
// (one 128-thread CTA per image; it reads the first / last 48 source bytes of each row, ~4 % of the
// pixels) and leaves their pass-2 values in a [n][H][6] scratch.  Then, per image (one CTA, 8 warps,
// ~108 KB shared memory, 2 CTAs per SM, <= 128 registers so the row chains keep their state in
// registers):
//   for each band of 192 rows:
//     F  8 warps : global (128-bit loads, three register sets in rotation, each pixel read once
//                  + 4 % halo) -> luma -> 2x2 rounded average -> u8 luma band in shared memory
//                  (the only copy of the plane)
//     C  8 warps : lane = row.  Horizontal 8-sums slide along the row in packed u16x2
//                  registers, vertical window sums come from warp shuffles, S2d -> float ->
//                  two-term-reciprocal division -> pass-3 chain; 64 samples per row go to a
//                  per-CTA L2-resident scratch (column-major, coalesced)
//   T  pass 4 over the scratch (64 column chains), decimate, then pdq_tail.cuh.
// HBM traffic is the pixels (read once) plus 36 B of results; the f32 planes of the reference
// never exist.
#include <stdlib.h>

#include "common.cuh"
#include "pdq_luma.cuh"
#include "pdq_tail.cuh"

namespace {

using namespace rh;

constexpr int FW = 512;           // plane width served by this kernel
constexpr int FLP = 528;          // luma row pitch in bytes: 512 + 16 zero bytes; 528 % 128 == 16
                                  // keeps the per-row LDS.128 of 8 consecutive lanes conflict-free
constexpr int FBAND = 192;        // output rows per band
constexpr int FTHREADS = 256;     // 8 warps: front end, row chains, tail
constexpr int FMAXL = FBAND + 7;  // luma rows per band including the vertical halo (window <= 8)
constexpr int P3_PITCH = 512;     // floats per column of the pass-3 scratch
constexpr size_t FSMEM = (size_t)FMAXL * FLP + 16 * DCT_PITCH * 4;   // luma band (aliased by the tail) + DCT matrix

static_assert(sizeof(TailSmem) <= (size_t)FMAXL * FLP, "tail scratch must fit in the luma band");

struct FusedArgs {
    const uint8_t *px;
    size_t row_pitch, img_pitch;
    int64_t n;
    int H;
    const float *p2e;  // [n][H][6] pass-2 values of the six inexact columns (pdq_edge_kernel)
    float *p3t;        // [gridDim.x][64][P3_PITCH]
    const float *dct;  // 16 x 64
    TailOut out;
    int64_t out_offset;
    int pf_mode;       // L2 prefetch: 0 = off, 1 = front inside the band, 2 = + band / image starts
    int pf_rows;       // plane rows between the prefetch front and the loads
    unsigned long long *phase_clk;   // nullptr, or [NPHASE] cycle totals of thread 0 of every CTA (RH_PDQ_PHASE_CLOCKS)
};

enum { PH_FRONT = 0, PH_CHAIN, PH_P4_STAGE, PH_P4_CHAIN, PH_TAIL, NPHASE };

// ------------------------------------------------------------------ front end ----

__device__ __forceinline__ void l2_prefetch_row(const uint8_t *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

constexpr int PF_ROWS = 16;   // plane rows between the L2 prefetch front and the loads

// Pull the source rows of plane rows [r0, r1) of the image at `base` into L2: one bulk prefetch per
// 3 KB source row, no registers, no shared memory.
template <int LAYOUT, bool DOWN2>
__device__ __forceinline__ void l2_prefetch_rows(const uint8_t *base, size_t row_pitch, int H, int r0, int r1,
                                                 int first, int step) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr uint32_t ROWB = 8 * SPP * CH * 64;
    for (int r = max(r0, 0) + first; r < min(r1, H); r += step) {
        const uint8_t *p = base + (size_t)(r * SPP) * row_pitch;
        l2_prefetch_row(p, ROWB);
        if (DOWN2) l2_prefetch_row(p + row_pitch, ROWB);
    }
}

// Phase F: fill the luma band.  64 threads per plane row (8 pixels each), 4 rows per sweep.  Each
// thread rotates through SETS register sets (3 for RGB / luma: 72 registers of pixels; 2 for RGBA):
// the loads of its rows k+1 .. k+SETS-1 are in flight while row k is converted (12 independent
// 128-bit loads per thread), so an L2-latency load has two conversions to land.  An L2 prefetch
// front runs PF_ROWS rows ahead inside the band (not across a row-chain phase: at ~4 TB/s a line
// survives only ~30 us in the 126 MB L2).
// Slots whose plane row lies outside the image (top of the first band, bottom of the last) are
// stored as zeros, and so are the 16 pad bytes of every row: the chain phase needs no clipping.
// PACKED: rows are contiguous (row_pitch == ROWB), so every address in the loop is the thread's
// pointer plus an immediate -- no pitch multiplies, no reloads of the pitch from the constant bank.
template <int LAYOUT, bool DOWN2, bool PACKED>
__device__ __forceinline__ void front_end(const uint8_t *__restrict__ src, size_t row_pitch_arg, int H, int Lr0, int nL,
                                          uint8_t *sL, int pf_mode, int pf_rows) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr int BYTES = 8 * SPP * CH;  // source bytes per thread and source row
    constexpr int NW = BYTES / 4;
    constexpr int SETS = NW * SPP * 3 <= 72 ? 3 : 2;
    constexpr uint32_t ROWB = BYTES * 64;
    const size_t row_pitch = PACKED ? (size_t)ROWB : row_pitch_arg;
    const int col8 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
    const int s_lo = max(0, -Lr0), s_hi = min(nL, H - Lr0);
    for (int s = rsub; s < nL; s += 4) {
        if (s < s_lo || s >= s_hi) *reinterpret_cast<uint2 *>(sL + (size_t)s * FLP + col8 * 8) = make_uint2(0u, 0u);
        if (col8 < 2) *reinterpret_cast<uint2 *>(sL + (size_t)s * FLP + FW + col8 * 8) = make_uint2(0u, 0u);
    }
    uint32_t w0[SETS][NW], w1[SETS][DOWN2 ? NW : 1];
    const size_t rstep = (size_t)(4 * SPP) * row_pitch;            // 4 plane rows further down
    int s = s_lo + rsub;
    const uint8_t *p = src + (size_t)((Lr0 + s) * SPP) * row_pitch + (size_t)col8 * BYTES;
    uint8_t *d = sL + (size_t)s * FLP + col8 * 8;
    const bool pf_on = PACKED && (threadIdx.x & 31) == 0 && pf_mode >= 1;
#pragma unroll
    for (int q = 0; q < SETS - 1; q++) {
        if (s + 4 * q < s_hi) {
            load_chunk<BYTES>(p + q * rstep, w0[q]);
            if (DOWN2) load_chunk<BYTES>(p + q * rstep + row_pitch, w1[q]);
        }
    }
    while (s < s_hi) {
#pragma unroll
        for (int q = 0; q < SETS; q++) {
            constexpr int AHEAD = SETS - 1;
            const int qa = (q + AHEAD) % SETS;          // the set converted AHEAD sweeps from now
            if (s + 4 * (q + AHEAD) < s_hi) {
                load_chunk<BYTES>(p + (q + AHEAD) * rstep, w0[qa]);
                if (DOWN2) load_chunk<BYTES>(p + (q + AHEAD) * rstep + row_pitch, w1[qa]);
            }
            if (q == 0 && pf_on) {
                // PACKED rows are contiguous in memory: the 4 SETS plane rows of the sweep group pf_rows
                // ahead are one byte range; lane 0 of each of the 8 warps prefetches an eighth of it
                const int f0 = s - rsub + pf_rows, f1 = min(f0 + 4 * SETS, s_hi);
                constexpr uint32_t PIECE = (uint32_t)(4 * SETS * SPP) * ROWB / 8;
                static_assert(PIECE % 16 == 0, "bulk prefetch granularity");
                const int beg = ((Lr0 + f0) * SPP) * (int)ROWB + (int)(threadIdx.x >> 5) * (int)PIECE;
                const int end = ((Lr0 + f1) * SPP) * (int)ROWB;
                if (beg < end) l2_prefetch_row(src + beg, (uint32_t)min((int)PIECE, end - beg));
            }
            if (s + 4 * q < s_hi) *reinterpret_cast<uint2 *>(d + 4 * q * FLP) = luma8<LAYOUT, DOWN2, NW>(w0[q], w1[q]);
        }
        s += 4 * SETS;
        p += SETS * rstep;
        d += 4 * SETS * FLP;
    }
}

// ---------------------------------------------------------------- edge columns ----

// Edge columns, step 1 (one thread per plane row, STRIDE threads): pass-1 values of the six columns
// whose clipped row window is 5, 6 or 7 wide (box_one_d_float phases 2 and 4, pdqhash.rs:372-378,
// :389-395) -- rounded quotients -- from the first and last 8-pixel chunk of each row, read
// straight from global memory.  Output is column-major: sE[c * EDGE_PITCH + row].
constexpr int EDGE_PITCH = 512 + 16;   // rows of a column + one walk batch of read-ahead, 16-byte aligned
template <int LAYOUT, bool DOWN2, int STRIDE>
__device__ __forceinline__ void edge_p1(const uint8_t *__restrict__ src, size_t row_pitch, int H, float *sE, int lane) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr int BYTES = 8 * SPP * CH;
    constexpr int NW = BYTES / 4;
    for (int lr = lane; lr < H; lr += STRIDE) {
        uint32_t l0[NW], l1[DOWN2 ? NW : 1], r0[NW], r1[DOWN2 ? NW : 1];
        const uint8_t *pl = src + (size_t)(lr * SPP) * row_pitch;
        const uint8_t *pr = pl + (size_t)63 * BYTES;
        load_chunk<BYTES>(pl, l0);
        load_chunk<BYTES>(pr, r0);
        if (DOWN2) {
            load_chunk<BYTES>(pl + row_pitch, l1);
            load_chunk<BYTES>(pr + row_pitch, r1);
        }
        const uint2 vl = luma8<LAYOUT, DOWN2, NW>(l0, l1);   // plane columns 0..7
        const uint2 vr = luma8<LAYOUT, DOWN2, NW>(r0, r1);   // plane columns 504..511
        // column 0: [0,4], column 1: [0,5], column 2: [0,6]
        const int s5 = (int)__dp4a(vl.x, 0x01010101u, 0u) + (int)(vl.y & 0xFFu);
        const int s6 = s5 + (int)((vl.y >> 8) & 0xFFu);
        const int s7 = s6 + (int)((vl.y >> 16) & 0xFFu);
        // column 510: [507,511], column 509: [506,511], column 508: [505,511]
        const int t5 = (int)__dp4a(vr.y, 0x01010101u, 0u) + (int)(vr.x >> 24);
        const int t6 = t5 + (int)((vr.x >> 16) & 0xFFu);
        const int t7 = t6 + (int)((vr.x >> 8) & 0xFFu);
        float *o = sE + lr;
        o[0 * EDGE_PITCH] = __fdiv_rn((float)s5, 5.0f);
        o[1 * EDGE_PITCH] = __fdiv_rn((float)s6, 6.0f);
        o[2 * EDGE_PITCH] = __fdiv_rn((float)s7, 7.0f);
        o[3 * EDGE_PITCH] = __fdiv_rn((float)t7, 7.0f);
        o[4 * EDGE_PITCH] = __fdiv_rn((float)t6, 6.0f);
        o[5 * EDGE_PITCH] = __fdiv_rn((float)t5, 5.0f);
    }
}

// ------------------------------------------------------------------ chain phase ----

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// sum over WC consecutive lanes starting at this lane (packed u16x2 fields cannot overflow:
// 8 rows x 2040 < 65536)
template <int WC>
__device__ __forceinline__ uint32_t window_sum_down(uint32_t h) {
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t a2 = h + __shfl_down_sync(FULL, h, 1);
    if (WC == 2) return a2;
    if (WC == 3) return a2 + __shfl_down_sync(FULL, h, 2);
    const uint32_t a4 = a2 + __shfl_down_sync(FULL, a2, 2);
    if (WC == 4) return a4;
    if (WC == 5) return a4 + __shfl_down_sync(FULL, h, 4);
    if (WC == 6) return a4 + __shfl_down_sync(FULL, a2, 4);
    if (WC == 7) return a4 + __shfl_down_sync(FULL, a2, 4) + __shfl_down_sync(FULL, h, 6);
    return a4 + __shfl_down_sync(FULL, a4, 4);
}

// RN(s / d) for the integer s held in the low (HI = false) or high 16-bit field of `packed` and
// d = 8 cnt (or 4 cnt), from a two-term reciprocal: yh = RN(1/d), yl = RN(RN(1 - d yh) yh) ~ 1/d - yh,
// q = fma(f, yh, RN(f yl)).  f yh + f yl is within 2^-45 of s / d, far inside the 1/6 ulp that
// separates s / d from a rounding tie, so q equals IEEE division for every (s <= 16320, cnt <= 8):
// tests/test_fused_model.py checks the whole set.  The conversion is one I2F.U16 with a half-word
// selector: it runs on the otherwise idle XU pipe instead of a PRMT + FADD pair on the busy ones.
template <bool HI>
__device__ __forceinline__ float div_exact(uint32_t packed, float yh, float yl) {
    const float f = __uint2float_rn(HI ? (packed >> 16) : (packed & 0xFFFFu));
    return __fmaf_rn(f, yh, __fmul_rn(f, yl));
}

struct Recip {
    float h, l;
};
__device__ __forceinline__ Recip recip2(float d) {
    Recip r;
    r.h = __frcp_rn(d);
    r.l = __fmul_rn(__fmaf_rn(-d, r.h, 1.0f), r.h);
    return r;
}

struct ChainState {
    uint32_t aprev;   // previous even-aligned entering pair (L[e+2], L[e+3])
    uint32_t Hp;      // (H[e-2], H[e-1]): horizontal clipped 8-sums, packed u16x2
    uint32_t S[4];    // entering sums of the last four pair steps (an 8-column delay line)
    float ring[8];    // P2 of the last 8 columns
    float sum;        // pass-3 running sum
};

enum { G_FIRST = 0, G_MID = 1, G_LAST = 2 };

// One 16-column group of the pass-3 row chain: entering columns e = 16 g - 4 .. 16 g + 11, two per
// step.  `cur` holds luma columns 16 g .. 16 g + 15 of the lane's row.
template <int WC, int KIND>
__device__ __forceinline__ void chain_group(ChainState &st, const uint4 cur, int g, Recip y8, Recip y4,
                                            const float *p2e, float *p3col, bool store) {
    const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
    for (int p = 0; p < 8; p++) {
        // entering bytes L[e+4], L[e+5] are bytes 2p, 2p+1 of this group
        const uint32_t aeven = prmt(cw[p >> 1], 0u, (p & 1) ? 0x4342u : 0x4140u);
        const uint32_t aodd = prmt(st.aprev, aeven, 0x5432u);  // (L[e+3], L[e+4])
        st.aprev = aeven;
        const uint32_t snew = aodd + aeven;
        st.Hp = st.Hp + snew - st.S[p & 3];                    // -> (H[e], H[e+1])
        st.S[p & 3] = snew;
        if (KIND == G_FIRST && p < 2) continue;                // e < 0: only the sliding sums advance
        if (KIND == G_LAST && p >= 2) continue;                // e > 511
        const uint32_t b = window_sum_down<WC>(st.Hp);         // S2d of this lane's output row
        float x0, x1;
        if (KIND == G_FIRST && p == 2) {                       // e = 0, 1: inexact columns
            x0 = p2e[0];
            x1 = p2e[1];
        } else if (KIND == G_FIRST && p == 3) {                // e = 2 inexact, e = 3 exact
            x0 = p2e[2];
            x1 = div_exact<true>(b, y8.h, y8.l);
        } else if (KIND == G_LAST && p == 0) {                 // e = 508, 509
            x0 = p2e[3];
            x1 = p2e[4];
        } else if (KIND == G_LAST && p == 1) {                 // e = 510 inexact; e = 511: 4-wide window
            x0 = p2e[5];
            x1 = div_exact<true>(b, y4.h, y4.l);
        } else {
            x0 = div_exact<false>(b, y8.h, y8.l);
            x1 = div_exact<true>(b, y8.h, y8.l);
        }
        // pass-3 chain (pdqhash.rs:366-387): entering column e, leaving e - 8, output column e - 4
        const bool full = !(KIND == G_FIRST && p < 6);         // e >= 8
        const int k0 = (2 * p) & 7, k1 = (2 * p + 1) & 7;
        st.sum = __fadd_rn(st.sum, x0);
        if (full) st.sum = __fsub_rn(st.sum, st.ring[k0]);
        st.ring[k0] = x0;
        if (full && (p == 2 || p == 6) && KIND != G_LAST) {
            // output column e - 4 = 8 j + 4 is decimation sample j (pdqhash.rs:439), j = (e - 8) / 8
            const int j = 2 * g - (p == 2 ? 1 : 0);
            if (store) __stcg(p3col + (size_t)j * P3_PITCH, __fmul_rn(st.sum, 0.125f));
        }
        st.sum = __fadd_rn(st.sum, x1);
        if (full) st.sum = __fsub_rn(st.sum, st.ring[k1]);
        st.ring[k1] = x1;
    }
}

// Phase C for one band: lane = row.  Warp w owns luma slots [w OPW, w OPW + 31] of the band window
// and produces output rows b0 + w OPW + lane for lane < OPW = 33 - WC.
template <int WC>
__device__ __forceinline__ void chain_phase(const uint8_t *sL, const float *p2e_img, float *p3t, int H, int b0,
                                            int rows_out, int nL) {
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1, OPW = 33 - WC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ro = warp * OPW + lane;   // output row within the band == luma slot of the window top
    const int r = b0 + ro;
    const bool store = lane < OPW && ro < rows_out;
    const int lo = max(0, r - HT), hi = min(H - 1, r + HB);
    const float cnt = (float)max(1, hi - lo + 1);   // rows in the clipped column window
    const Recip y8 = recip2(8.0f * cnt), y4 = recip2(4.0f * cnt);
    const uint8_t *rowp = sL + (size_t)min(ro, nL - 1) * FLP;
    const float *p2e = p2e_img + (size_t)min(r, H - 1) * 6;   // edge-column values of plane row r (pdq_edge_kernel)
    float *p3col = p3t + r;
    ChainState st;
    st.aprev = 0u;
    st.Hp = 0u;
    st.sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) st.S[i] = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) st.ring[i] = 0.0f;
    // the 16 luma bytes of group g + 1 are fetched while group g runs (no LDS latency on the chain)
    uint4 cur = *reinterpret_cast<const uint4 *>(rowp);
    uint4 nxt = *reinterpret_cast<const uint4 *>(rowp + 16);
    chain_group<WC, G_FIRST>(st, cur, 0, y8, y4, p2e, p3col, store);
#pragma unroll 1
    for (int g = 1; g < 32; g++) {
        cur = nxt;
        nxt = *reinterpret_cast<const uint4 *>(rowp + 16 * g + 16);   // g = 31: the zero pad at column 512
        chain_group<WC, G_MID>(st, cur, g, y8, y4, p2e, p3col, store);
    }
    chain_group<WC, G_LAST>(st, nxt, 32, y8, y4, p2e, p3col, store);
    // first output of the shrink phase: column 508 = sample 63, window of 7 (pdqhash.rs:389-395).
    // P2[504] entered at g = 31, p = 6 and sits in ring[(2*6) & 7].
    st.sum = __fsub_rn(st.sum, st.ring[4]);
    if (store) __stcg(p3col + (size_t)63 * P3_PITCH, __fdiv_rn(st.sum, 7.0f));
}

// ------------------------------------------------------------------------ tail ----

// Pass 4: the column chains over the pass-3 samples (window WC, length H) for the 64 decimated
// columns, keeping the 64 decimated rows (pdqhash.rs:435).  The slab is pulled from L2 into shared
// memory P4_ROWS rows at a time with cp.async (16 bytes per request, no registers, every request of a
// chunk in flight at once) into two buffers: chunk c + 1 lands while threads 0..63 (one per column)
// walk chunk c at shared-memory latency, so only the first chunk's L2 round trip is exposed.
// The pitch is a multiple of 4 floats with pitch / 4 odd: the 128-bit accesses of the walk (8 lanes
// per wavefront) are conflict-free.
constexpr int P4_ROWS = 128;
constexpr int P4_PITCH = P4_ROWS + 4;
static_assert((P4_PITCH / 4) % 2 == 1 && P4_PITCH % 4 == 0, "pass-4 staging pitch");
static_assert((64 * (P4_ROWS / 4)) % FTHREADS == 0 && P4_ROWS % 8 == 0, "pass-4 staging has no remainder");
constexpr size_t P4_STAGE_OFF = (sizeof(TailSmem) + 15) & ~size_t(15);   // 16-byte aligned for the 128-bit accesses
static_assert(P4_STAGE_OFF + 2 * 64 * P4_PITCH * 4 <= (size_t)FMAXL * FLP, "pass-4 staging must fit beside the tail scratch");

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
__device__ __forceinline__ void p4_issue(const float *p3t, int c0, float *stg) {
#pragma unroll 1
    for (int idx = threadIdx.x; idx < 64 * (P4_ROWS / 4); idx += FTHREADS) {
        const int col = idx / (P4_ROWS / 4), q = idx % (P4_ROWS / 4);
        cp_async16(stg + col * P4_PITCH + 4 * q, p3t + (size_t)col * P3_PITCH + min(c0 + 4 * q, P3_PITCH - 4));
    }
    cp_async_commit();
}

template <int WC>
__device__ __forceinline__ void pass4(const float *p3t, int H, float *B, float *stage, float *shr, PhaseClock &clk) {
    constexpr int HALF = (WC + 2) / 2, HB = HALF - 1;
    const int j = threadIdx.x;
    float sum = 0.0f;
    float prev[8];
#pragma unroll
    for (int k = 0; k < 8; k++) prev[k] = 0.0f;
    p4_issue(p3t, 0, stage);
    if (P4_ROWS < H) p4_issue(p3t, P4_ROWS, stage + 64 * P4_PITCH);
    int buf = 0;
    for (int c0 = 0; c0 < H; c0 += P4_ROWS, buf ^= 1) {
        float *stg = stage + buf * (64 * P4_PITCH);
        const int rows = min(P4_ROWS, H - c0);
        const bool last = c0 + P4_ROWS >= H;
        // the rows that leave during the shrink phase (H-WC .. H-WC+HB-1), fetched under the walk
        float leave[HB > 0 ? HB : 1];
        if (last && j < 64) {
#pragma unroll
            for (int k = 0; k < HB; k++) leave[k] = __ldcg(p3t + (size_t)j * P3_PITCH + (H - WC + k));
        }
        if (last)
            cp_async_wait<0>();
        else
            cp_async_wait<1>();   // everything but the chunk after this one has landed
        __syncthreads();
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
        }
        __syncthreads();
        clk.lap(PH_P4_CHAIN);
        p4_gather<WC>(stg, shr, H, c0, rows, last, B);
        if (c0 + 2 * P4_ROWS < H) {   // this buffer takes the chunk after the next one
            __syncthreads();
            p4_issue(p3t, c0 + 2 * P4_ROWS, stg);
        }
    }
}

template <int LAYOUT, bool DOWN2, int WC, bool PACKED>
__global__ void __launch_bounds__(FTHREADS, 2) pdq_fused_kernel(const FusedArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *sL = smem;
    TailSmem &ts = *reinterpret_cast<TailSmem *>(smem);   // aliases the luma band, used after the last band
    float *sD = reinterpret_cast<float *>(smem + (size_t)FMAXL * FLP);   // DCT matrix, resident for the whole kernel
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, OPW = 33 - WC;
    constexpr int NWC = (FBAND + OPW - 1) / OPW;           // warps that run row chains
    static_assert(NWC <= FTHREADS / 32, "one warp per OPW output rows of a band");
    const int H = a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *p3t = a.p3t + (size_t)blockIdx.x * 64 * P3_PITCH;
    for (int idx = threadIdx.x; idx < 1024; idx += FTHREADS) sD[(idx >> 6) * DCT_PITCH + (idx & 63)] = a.dct[idx];
    PhaseClock clk;
    clk.start(a.phase_clk);

    for (int64_t img = blockIdx.x; img < a.n; img += gridDim.x) {
        const uint8_t *src = a.px + (size_t)img * a.img_pitch;
        const uint8_t *next_src = img + gridDim.x < a.n ? src + (size_t)gridDim.x * a.img_pitch : nullptr;
        for (int b0 = 0; b0 < H; b0 += FBAND) {
            const int rows_out = min(FBAND, H - b0);
            const int Lr0 = b0 - HT;
            const int nL = rows_out + WC - 1;
            front_end<LAYOUT, DOWN2, PACKED>(src, a.row_pitch, H, Lr0, nL, sL, a.pf_mode, a.pf_rows);
            __syncthreads();
            clk.lap(PH_FRONT);
            if (warp < NWC) chain_phase<WC>(sL, a.p2e + (size_t)img * H * 6, p3t, H, b0, rows_out, nL);
            // warm L2 with the first PF_ROWS rows of whatever the front end loads next (the next band
            // of this image, else the first band of the CTA's next image), a few us before it starts
            if (lane == 0 && a.pf_mode >= 2) {
                if (b0 + FBAND < H)
                    l2_prefetch_rows<LAYOUT, DOWN2>(src, a.row_pitch, H, b0 + FBAND - HT, b0 + FBAND - HT + a.pf_rows, warp, 8);
                else if (next_src != nullptr)
                    l2_prefetch_rows<LAYOUT, DOWN2>(next_src, a.row_pitch, H, 0, a.pf_rows, warp, 8);
            }
            __syncthreads();
            clk.lap(PH_CHAIN);
        }
        // pass 4 + decimation into the tail's 64 x 64 buffer, then quality / DCT / hash
        pass4<WC>(p3t, H, ts.B, reinterpret_cast<float *>(smem + P4_STAGE_OFF), ts.T, clk);
        __syncthreads();
        clk.lap(PH_P4_STAGE);   // (the last gather)
        const size_t oimg = (size_t)img + (size_t)a.out_offset;
        const float q = tail_quality(ts);
        if (threadIdx.x == 0 && a.out.quality) a.out.quality[oimg] = q;
        tail_dct(ts, sD);
        if (a.out.coeffs) a.out.coeffs[oimg * 256 + threadIdx.x] = ts.C[threadIdx.x];
        tail_hashes(ts, a.out, oimg);
        __syncthreads();   // the next image's front end overwrites the aliased tail scratch
        clk.lap(PH_TAIL);
    }
}

// The six inexact columns for a whole chunk of images, one CTA of EDGE_THREADS threads per image:
// edge_p1 over all rows (one thread per row), then the column pass of box_one_d_float (pdqhash.rs:
// 341-396) on each of the six columns -- the same in-place walk pass 4 uses, one lane per column, plus
// the shrink phase -- then the division by the clipped window size, results to p2e[img][row][6].
// It reads the first and last 48 bytes of every source row (~4 % of the pixels, 32-byte sectors).
constexpr int EDGE_THREADS = 128;

template <int LAYOUT, bool DOWN2, int WC>
__global__ void __launch_bounds__(EDGE_THREADS) pdq_edge_kernel(const uint8_t *__restrict__ px, size_t row_pitch, size_t img_pitch,
                                                                int64_t n, int H, float *__restrict__ p2e) {
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1;
    __shared__ __align__(16) float sE[6 * EDGE_PITCH];
    __shared__ float shr[6 * 4];   // the shrink-phase sums
    const int t = threadIdx.x;
    for (int64_t img = blockIdx.x; img < n; img += gridDim.x) {
        edge_p1<LAYOUT, DOWN2, EDGE_THREADS>(px + (size_t)img * img_pitch, row_pitch, H, sE, t);
        __syncthreads();
        if (t < 6) {
            float *col = sE + t * EDGE_PITCH;
            float leave[HB > 0 ? HB : 1];
#pragma unroll
            for (int k = 0; k < HB; k++) leave[k] = col[H - WC + k];   // the walk overwrites them
            float sum = 0.0f, prev[8];
#pragma unroll
            for (int k = 0; k < 8; k++) prev[k] = 0.0f;
            p4_walk<WC>(col, 0, H, sum, prev);   // col[i] = window sum after row i has entered
            sum = col[H - 1];
#pragma unroll
            for (int k = 0; k < HB; k++) {
                sum = __fsub_rn(sum, leave[k]);
                shr[t * 4 + k] = sum;
            }
        }
        __syncthreads();
        float *out = p2e + (size_t)img * H * 6;
        for (int idx = t; idx < H * 6; idx += EDGE_THREADS) {
            const int o = idx / 6, c = idx - 6 * o;
            const int cnt = min(H - 1, o + HB) - max(0, o - HT) + 1;   // pdqhash.rs:375, :383, :392
            const float v = o >= H - HB ? shr[c * 4 + (o - (H - HB))] : sE[c * EDGE_PITCH + o + HB];
            out[idx] = __fdiv_rn(v, (float)cnt);
        }
        __syncthreads();
    }
}

template <int LAYOUT, bool DOWN2, int WC>
int launch_fused(rh_ctx *ctx, const FusedArgs &a, int grid) {
    pdq_edge_kernel<LAYOUT, DOWN2, WC><<<(unsigned)(a.n < 16 * 148 * 4 ? a.n : 16 * 148 * 4), EDGE_THREADS, 0, ctx->stream>>>(
        a.px, a.row_pitch, a.img_pitch, a.n, a.H, const_cast<float *>(a.p2e));
    RH_LAUNCHED(ctx, "pdq_edge_kernel");
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    const bool packed = a.row_pitch == (size_t)FW * (DOWN2 ? 2 : 1) * CH;
    auto kern = packed ? pdq_fused_kernel<LAYOUT, DOWN2, WC, true> : pdq_fused_kernel<LAYOUT, DOWN2, WC, false>;
    RH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FSMEM));
    kern<<<grid, FTHREADS, FSMEM, ctx->stream>>>(a);
    RH_LAUNCHED(ctx, "pdq_fused_kernel");
    return RH_OK;
}

template <int LAYOUT, bool DOWN2>
int dispatch_wc(rh_ctx *ctx, const FusedArgs &a, int grid, int wc) {
    if (wc == 4) return launch_fused<LAYOUT, DOWN2, 4>(ctx, a, grid);
    if (wc == 5) return launch_fused<LAYOUT, DOWN2, 5>(ctx, a, grid);
    if (wc == 6) return launch_fused<LAYOUT, DOWN2, 6>(ctx, a, grid);
    if (wc == 7) return launch_fused<LAYOUT, DOWN2, 7>(ctx, a, grid);
    if (wc == 8) return launch_fused<LAYOUT, DOWN2, 8>(ctx, a, grid);
    return fail(ctx, RH_EUNSUPPORTED, "fused PDQ kernel: column window not instantiated");
}

}  // namespace

namespace rh {

// planes 512 wide whose column window ceil(H / 64) is 4 .. 8 (H in 193..512: every landscape shape from
// 8:3 to 1:1 after the reference's resize to 512 wide)
int pdq_fused_supported(int W, int H) {
    if (W != FW || H > 512) return 0;
    const int wc = (H + 63) / 64;
    return wc >= 4 && wc <= 8;
}

// 128-bit loads need 16-byte aligned rows
int pdq_fused_aligned(const void *px, size_t row_pitch, size_t img_pitch) {
    return ((reinterpret_cast<uintptr_t>(px) | row_pitch | img_pitch) & 15) == 0;
}

int pdq_fused_run(rh_ctx *ctx, const uint8_t *d_px, int layout, bool down2, int64_t n, int W, int H, size_t row_pitch,
                  size_t img_pitch, const TailOut &out, int64_t out_offset, const float *d_dct) {
    if (!pdq_fused_supported(W, H)) return fail(ctx, RH_EUNSUPPORTED, "fused PDQ kernel: unsupported plane size");
    if ((reinterpret_cast<uintptr_t>(d_px) | row_pitch | img_pitch) & 15)
        return fail(ctx, RH_EINVAL, "fused PDQ kernel: pixels must be 16-byte aligned (pdq_fused_aligned)");
    int grid = ctx->sm_count * 2;
    if (grid > n) grid = (int)n;
    void *p, *p_p3t, *p_p2e;
    RH_TRY(scratch(ctx, S_W3, (size_t)grid * 64 * P3_PITCH * sizeof(float), &p_p3t));
    RH_TRY(scratch(ctx, S_W4, (size_t)n * H * 6 * sizeof(float), &p_p2e));
    FusedArgs a;
    a.p2e = (const float *)p_p2e;
    // rh_ctx_set_option("pdq.phase_clocks", 1): per-phase cycle totals of thread 0 of every CTA, printed after the kernel
    const bool clocks = ctx->pdq_phase_clocks != 0;
    a.phase_clk = nullptr;
    if (clocks) {
        RH_TRY(scratch(ctx, S_W8, NPHASE * sizeof(unsigned long long), &p));
        a.phase_clk = (unsigned long long *)p;
        RH_CUDA(ctx, cudaMemsetAsync(p, 0, NPHASE * sizeof(unsigned long long), ctx->stream));
    }
    a.px = d_px;
    a.row_pitch = row_pitch;
    a.img_pitch = img_pitch;
    a.n = n;
    a.H = H;
    a.p3t = (float *)p_p3t;
    a.dct = d_dct;
    a.out = out;
    a.out_offset = out_offset;
    a.pf_mode = ctx->pdq_prefetch;
    a.pf_rows = ctx->pdq_prefetch_rows;
    const int wc = (H + 63) / 64;
    int rc;
    if (layout == RH_LAYOUT_RGB8)
        rc = down2 ? dispatch_wc<RH_LAYOUT_RGB8, true>(ctx, a, grid, wc) : dispatch_wc<RH_LAYOUT_RGB8, false>(ctx, a, grid, wc);
    else if (layout == RH_LAYOUT_RGBA8)
        rc = down2 ? dispatch_wc<RH_LAYOUT_RGBA8, true>(ctx, a, grid, wc) : dispatch_wc<RH_LAYOUT_RGBA8, false>(ctx, a, grid, wc);
    else
        rc = down2 ? dispatch_wc<RH_LAYOUT_LUMA8, true>(ctx, a, grid, wc) : dispatch_wc<RH_LAYOUT_LUMA8, false>(ctx, a, grid, wc);
    if (rc == RH_OK && clocks) {
        unsigned long long h[NPHASE];
        RH_CUDA(ctx, cudaMemcpyAsync(h, a.phase_clk, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        static const char *names[NPHASE] = {"front", "chain", "p4_stage", "p4_chain", "tail"};
        unsigned long long tot = 0;
        for (int i = 0; i < NPHASE; i++) tot += h[i];
        fprintf(stderr, "[pdq_fused phases] n=%lld", (long long)n);
        for (int i = 0; i < NPHASE; i++)
            fprintf(stderr, "  %s %.0f cyc/img (%.1f%%)", names[i], (double)h[i] / (double)n, 100.0 * (double)h[i] / (double)tot);
        fprintf(stderr, "\n");
    }
    return rc;
}

}  // namespace rh
