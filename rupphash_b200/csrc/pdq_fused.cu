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
// Data flow, per image (one CTA, 8 warps, 114 KB shared memory, 2 CTAs per SM, <= 128 registers so the
// row chains keep their state in registers):
//   for each band of 192 rows:
//     F  8 warps : global (128-bit evict-first loads, three register sets in rotation, an L2 prefetch
//                  front through the bulk-copy engine; each pixel read ONCE: the rows two bands share
//                  are carried over in shared memory) -> luma -> 2x2 rounded average -> u8 luma band in
//                  shared memory (the only copy of the plane).  The two CTAs of an SM take turns here.
//     E  the six inexact columns from the band's first / last eight luma bytes per row: pass-1
//                  quotients (one thread per row), then the reference's column chain on six lanes,
//                  in place, continued from band to band in registers
//     C  8 warps : lane = row.  Horizontal 8-sums slide along the row in packed u16x2
//                  registers, vertical window sums come from warp shuffles, S2d -> float ->
//                  two-term-reciprocal division -> pass-3 chain; 64 samples per row go to a
//                  per-CTA L2-resident slab (column-major, coalesced, evict-last)
//   P  pass 4 over the slab (pdq_pass4.cuh: chunks staged by cp.async.bulk + mbarrier, 64 column chains),
//      decimate, then T: pdq_tail.cuh.
// HBM traffic is the pixels (read once) plus 36 B of results: 1.03x algorithmic by ncu; the f32 planes
// of the reference never exist.
#include <stdlib.h>

#include "common.cuh"
#include "pdq_pass4.cuh"
#include "tma.cuh"

namespace {

using namespace rh;

constexpr int FW = 512;           // plane width served by this kernel
constexpr int FLP = 528;          // luma row pitch in bytes: 512 + 16 zero bytes; 528 % 128 == 16
                                  // keeps the per-row LDS.128 of 8 consecutive lanes conflict-free
constexpr int FBAND = 192;        // output rows per band
constexpr int FMAXL = FBAND + 7;  // luma rows per band including the vertical halo (window <= 8)
constexpr int E_PITCH = FMAXL + 5;   // entries per edge column: index e = luma slot + 1 (e = 0: the row above the band)
constexpr size_t FSMEM = (size_t)FMAXL * FLP + 16 * DCT_PITCH * 4 + 6 * E_PITCH * 4;   // luma band (aliased by the tail) + DCT matrix + edge columns
static_assert(2 * (FSMEM + 1024) <= 233472, "two CTAs per SM");

static_assert(P4_SMEM_BYTES <= (size_t)FMAXL * FLP, "tail scratch + pass-4 staging must fit in the luma band");

struct FusedArgs {
    const uint8_t *px;
    size_t row_pitch, img_pitch;
    int64_t n;
    int H;
    float *p3t;        // [gridDim.x][64][P3_PITCH]
    const float *dct;  // 16 x 64
    TailOut out;
    int64_t out_offset;
    int pf_mode;       // L2 prefetch: 0 = off, 1 = front inside the band, 2 = + band / image starts
    int pf_rows;       // plane rows between the prefetch front and the loads
    int variant;       // A-B switches (rh_ctx_set_option "pdq.variant"), each bit turns one default OFF:
                       // 1 = pixels evict-first, 2 = slab evict-last, 4 = discard the slab's L2 lines after
                       // pass 4, 8 = the two CTAs of an SM alternate in the load phase
    unsigned long long *phase_clk;   // nullptr, or [NPHASE] cycle totals of thread 0 of every CTA (ctx option "pdq.phase_clocks")
};


// ------------------------------------------------------------------ front end ----

__device__ __forceinline__ void l2_prefetch_row(const uint8_t *p, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(pol) : "memory");
}

constexpr int PF_ROWS = 32;   // plane rows between the L2 prefetch front and the loads (ctx default; 16..32 measured equal to 2 % better)

// Pull the source rows of plane rows [r0, r1) of the image at `base` into L2: one bulk prefetch per
// 3 KB source row, no registers, no shared memory.
template <int LAYOUT, bool DOWN2>
__device__ __forceinline__ void l2_prefetch_rows(const uint8_t *base, size_t row_pitch, int H, int r0, int r1,
                                                 int first, int step, uint64_t pol) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr uint32_t ROWB = 8 * SPP * CH * 64;
    for (int r = max(r0, 0) + first; r < min(r1, H); r += step) {
        const uint8_t *p = base + (size_t)(r * SPP) * row_pitch;
        l2_prefetch_row(p, ROWB, pol);
        if (DOWN2) l2_prefetch_row(p + row_pitch, ROWB, pol);
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
                                          int s_begin, uint8_t *sL, int pf_mode, int pf_rows, uint64_t pol) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr int BYTES = 8 * SPP * CH;  // source bytes per thread and source row
    constexpr int NW = BYTES / 4;
    // as many sets as 72 registers of pixels hold, at most 8: without the 2x pre-downsample a set is only
    // 24 bytes (RGB), and 3 sets left the phase waiting on latency (512 x 512: 777 cycles per 6 KB sweep)
    constexpr int SETS = 72 / (NW * SPP) >= 8 ? 8 : (72 / (NW * SPP) >= 3 ? 72 / (NW * SPP) : 2);
    constexpr uint32_t ROWB = BYTES * 64;
    const size_t row_pitch = PACKED ? (size_t)ROWB : row_pitch_arg;
    const int col8 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
    // slots below s_begin were carried over from the previous band (rows it had already converted)
    const int s_lo = max(s_begin, -Lr0), s_hi = min(nL, H - Lr0);
    for (int s = s_begin + rsub; s < nL; s += 4) {
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
            load_px<BYTES>(p + q * rstep, w0[q], pol);
            if (DOWN2) load_px<BYTES>(p + q * rstep + row_pitch, w1[q], pol);
        }
    }
    while (s < s_hi) {
#pragma unroll
        for (int q = 0; q < SETS; q++) {
            constexpr int AHEAD = SETS - 1;
            const int qa = (q + AHEAD) % SETS;          // the set converted AHEAD sweeps from now
            if (s + 4 * (q + AHEAD) < s_hi) {
                load_px<BYTES>(p + (q + AHEAD) * rstep, w0[qa], pol);
                if (DOWN2) load_px<BYTES>(p + (q + AHEAD) * rstep + row_pitch, w1[qa], pol);
            }
            if (q == 0 && pf_on) {
                // PACKED rows are contiguous in memory: the 4 SETS plane rows of the sweep group pf_rows
                // ahead are one byte range; lane 0 of each of the 8 warps prefetches an eighth of it
                // (the front stays ahead of the loads, which run 4 (SETS - 1) rows ahead of the conversions)
                const int f0 = s - rsub + max(pf_rows, 4 * (SETS - 1) + 16), f1 = min(f0 + 4 * SETS, s_hi);
                constexpr uint32_t PIECE = (uint32_t)(4 * SETS * SPP) * ROWB / 8;
                static_assert(PIECE % 16 == 0, "bulk prefetch granularity");
                const int beg = ((Lr0 + f0) * SPP) * (int)ROWB + (int)(threadIdx.x >> 5) * (int)PIECE;
                const int end = ((Lr0 + f1) * SPP) * (int)ROWB;
                if (beg < end) l2_prefetch_row(src + beg, (uint32_t)min((int)PIECE, end - beg), pol);
            }
            if (s + 4 * q < s_hi) {
                RH_CHECK_IDX(s + 4 * q, FMAXL);
                *reinterpret_cast<uint2 *>(d + 4 * q * FLP) = luma8<LAYOUT, DOWN2, NW>(w0[q], w1[q]);
            }
        }
        s += 4 * SETS;
        p += SETS * rstep;
        d += 4 * SETS * FLP;
    }
}

// ---------------------------------------------------------------- edge columns ----

// The six plane columns 0,1,2,508,509,510 have clipped row windows of 5, 6 and 7 pixels, so their
// pass-1 values are rounded quotients (box_one_d_float phases 2 and 4, pdqhash.rs:372-378, :389-395)
// and their pass-2 values need the reference's real sequential column chain.  Both are done inside
// the fused kernel from the luma band (no second read of the pixels):
//   edge_inputs  one thread per luma row of the band: the six pass-1 quotients -> sE[c][slot + 1]
//   edge_walk    six lanes, one per column: the column pass of box_one_d_float (pdqhash.rs:341-396)
//                continued from band to band in registers, in place: on return sE[c][r - b0] is the
//                running window sum of output row r (the row chains divide it by their row count)
// Entry e = slot + 1 of a column belongs to plane row Lr0 + slot; entry 0 is the row above the band
// (carried over from the previous band like the luma halo).
// RN(s / d) for the small integer sums s <= 7 * 255 and d = 5, 6, 7 from a two-term reciprocal (the same
// construction as div_exact below; tests/test_fused_model.py checks every (s, d) in exact arithmetic)
__device__ __forceinline__ float div_small(int s, float yh, float yl) {
    const float f = (float)s;
    return __fmaf_rn(f, yh, __fmul_rn(f, yl));
}

__device__ __forceinline__ void edge_inputs(const uint8_t *sL, float *sE, int s0, int s1) {
    // yh = RN(1/d), yl = RN(RN(1 - d yh) yh) for d = 5, 6, 7
    const float h5 = __frcp_rn(5.0f), l5 = __fmul_rn(__fmaf_rn(-5.0f, h5, 1.0f), h5);
    const float h6 = __frcp_rn(6.0f), l6 = __fmul_rn(__fmaf_rn(-6.0f, h6, 1.0f), h6);
    const float h7 = __frcp_rn(7.0f), l7 = __fmul_rn(__fmaf_rn(-7.0f, h7, 1.0f), h7);
    for (int s = s0 + (int)threadIdx.x; s < s1; s += FTHREADS) {
        const uint8_t *row = sL + (size_t)s * FLP;
        const uint2 vl = *reinterpret_cast<const uint2 *>(row);              // plane columns 0..7
        const uint2 vr = *reinterpret_cast<const uint2 *>(row + FW - 8);     // plane columns 504..511
        // column 0: [0,4], column 1: [0,5], column 2: [0,6]
        const int s5 = (int)__dp4a(vl.x, 0x01010101u, 0u) + (int)(vl.y & 0xFFu);
        const int s6 = s5 + (int)((vl.y >> 8) & 0xFFu);
        const int s7 = s6 + (int)((vl.y >> 16) & 0xFFu);
        // column 510: [507,511], column 509: [506,511], column 508: [505,511]
        const int t5 = (int)__dp4a(vr.y, 0x01010101u, 0u) + (int)(vr.x >> 24);
        const int t6 = t5 + (int)((vr.x >> 16) & 0xFFu);
        const int t7 = t6 + (int)((vr.x >> 8) & 0xFFu);
        float *o = sE + s + 1;
        RH_CHECK_IDX(s + 1, E_PITCH);
        o[0 * E_PITCH] = div_small(s5, h5, l5);
        o[1 * E_PITCH] = div_small(s6, h6, l6);
        o[2 * E_PITCH] = div_small(s7, h7, l7);
        o[3 * E_PITCH] = div_small(t7, h7, l7);
        o[4 * E_PITCH] = div_small(t6, h6, l6);
        o[5 * E_PITCH] = div_small(t5, h5, l5);
    }
}

// `sum` is the window sum after the last row that entered (or left) in the previous band (0 before band 0).
template <int WC>
__device__ __forceinline__ void edge_walk(float *col, int H, int b0, int rows_out, float &sum) {
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1;
    const int Lr0 = b0 - HT;
    const int i_first = b0 == 0 ? 0 : b0 + HB, i_last = min(b0 + rows_out - 1 + HB, H - 1);
    // entering row i sits at e = i - Lr0 + 1, the row that leaves (i - WC) at e - WC, and the window
    // sum of output row i - HB is written over that leaving entry (position i - HB - b0)
    float *p = col + (i_first - Lr0 + 1);
    int i = i_first;
    for (; i < WC && i <= i_last; i++, p++) {      // the window is still filling (pdqhash.rs:366-378)
        sum = __fadd_rn(sum, p[0]);
        if (i >= HB) p[-WC] = sum;
    }
    // steady state (pdqhash.rs:380-387), four rows per step: the operands of the next four rows are
    // fetched before the current four are summed, so only the add / subtract pairs sit on the chain
    // (reading up to seven entries past the last row stays inside the column: E_PITCH leaves room)
    float x[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = p[k];
        o[k] = p[k - WC];
    }
    for (; i + 3 <= i_last; i += 4, p += 4) {
        float xn[4], on[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            xn[k] = p[4 + k];
            on[k] = p[4 + k - WC];
        }
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sum = __fsub_rn(__fadd_rn(sum, x[k]), o[k]);
            r[k] = sum;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            p[k - WC] = r[k];
            x[k] = xn[k];
            o[k] = on[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (i + k <= i_last) {
            sum = __fsub_rn(__fadd_rn(sum, x[k]), o[k]);
            p[k - WC] = sum;
        }
    }
    // shrink phase (pdqhash.rs:389-395): the outputs H-HB .. H-1 that belong to this band; the row that
    // leaves for output row r sits at position r - b0, which is also where the sum goes
    for (int r = max(H - HB, b0); r < b0 + rows_out; r++) {
        float *q = col + (r - b0);
        sum = __fsub_rn(sum, q[0]);
        q[0] = sum;
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
                                            const float *pe, float cnt, float *p3col, bool store, uint64_t pol_slab) {
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
            x0 = __fdiv_rn(pe[0 * E_PITCH], cnt);     // pdqhash.rs:375, :383, :392: running sum / running count
            x1 = __fdiv_rn(pe[1 * E_PITCH], cnt);
        } else if (KIND == G_FIRST && p == 3) {                // e = 2 inexact, e = 3 exact
            x0 = __fdiv_rn(pe[2 * E_PITCH], cnt);
            x1 = div_exact<true>(b, y8.h, y8.l);
        } else if (KIND == G_LAST && p == 0) {                 // e = 508, 509
            x0 = __fdiv_rn(pe[3 * E_PITCH], cnt);
            x1 = __fdiv_rn(pe[4 * E_PITCH], cnt);
        } else if (KIND == G_LAST && p == 1) {                 // e = 510 inexact; e = 511: 4-wide window
            x0 = __fdiv_rn(pe[5 * E_PITCH], cnt);
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
            if (store) st_slab(p3col + (size_t)j * P3_PITCH, __fmul_rn(st.sum, 0.125f), pol_slab);
        }
        st.sum = __fadd_rn(st.sum, x1);
        if (full) st.sum = __fsub_rn(st.sum, st.ring[k1]);
        st.ring[k1] = x1;
    }
}

// Phase C for one band: lane = row.  Warp w owns luma slots [w OPW, w OPW + 31] of the band window
// and produces output rows b0 + w OPW + lane for lane < OPW = 33 - WC.
template <int WC>
__device__ __forceinline__ void chain_phase(const uint8_t *sL, const float *sE, float *p3t, int H, int b0,
                                            int rows_out, int nL, uint64_t pol_slab) {
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1, OPW = 33 - WC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ro = warp * OPW + lane;   // output row within the band == luma slot of the window top
    const int r = b0 + ro;
    const bool store = lane < OPW && ro < rows_out;
    const int lo = max(0, r - HT), hi = min(H - 1, r + HB);
    const float cnt = (float)max(1, hi - lo + 1);   // rows in the clipped column window
    const Recip y8 = recip2(8.0f * cnt), y4 = recip2(4.0f * cnt);
    RH_CHECK_IDX(min(ro, nL - 1), FMAXL);
    if (store) RH_CHECK_IDX(r, P3_PITCH);
    const uint8_t *rowp = sL + (size_t)min(ro, nL - 1) * FLP;
    const float *pe = sE + min(ro, rows_out - 1);   // window sums of the six edge columns for this lane's row (edge_walk)
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
    chain_group<WC, G_FIRST>(st, cur, 0, y8, y4, pe, cnt, p3col, store, pol_slab);
#pragma unroll 4
    for (int g = 1; g < 32; g++) {
        cur = nxt;
        nxt = *reinterpret_cast<const uint4 *>(rowp + 16 * g + 16);   // g = 31: the zero pad at column 512
        chain_group<WC, G_MID>(st, cur, g, y8, y4, pe, cnt, p3col, store, pol_slab);
    }
    chain_group<WC, G_LAST>(st, nxt, 32, y8, y4, pe, cnt, p3col, store, pol_slab);
    // first output of the shrink phase: column 508 = sample 63, window of 7 (pdqhash.rs:389-395).
    // P2[504] entered at g = 31, p = 6 and sits in ring[(2*6) & 7].
    st.sum = __fsub_rn(st.sum, st.ring[4]);
    if (store) st_slab(p3col + (size_t)63 * P3_PITCH, __fdiv_rn(st.sum, 7.0f), pol_slab);
}

__device__ unsigned int g_front_lock[512];   // one per SM: which of its two CTAs may run the load phase

template <int LAYOUT, bool DOWN2, int WC, bool PACKED>
__global__ void __launch_bounds__(FTHREADS, 2) pdq_fused_kernel(const FusedArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *sL = smem;
    TailSmem &ts = *reinterpret_cast<TailSmem *>(smem);   // aliases the luma band, used after the last band
    float *sD = reinterpret_cast<float *>(smem + (size_t)FMAXL * FLP);   // DCT matrix, resident for the whole kernel
    float *sE = sD + 16 * DCT_PITCH;                                     // the six edge columns of the current band
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1, OPW = 33 - WC;
    constexpr int NWC = (FBAND + OPW - 1) / OPW;           // warps that run row chains
    static_assert(NWC <= FTHREADS / 32, "one warp per OPW output rows of a band");
    const int H = a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *p3t = a.p3t + (size_t)blockIdx.x * 64 * P3_PITCH;
    for (int idx = threadIdx.x; idx < 1024; idx += FTHREADS) sD[(idx >> 6) * DCT_PITCH + (idx & 63)] = a.dct[idx];
    PhaseClock clk;
    clk.start(a.phase_clk);
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    // L2 policies (a.variant bits 0 / 1 switch them off for A-B runs): streaming pixels evict-first, slab evict-last
    const uint64_t pol_px = (a.variant & 1) ? l2_policy_evict_normal() : l2_policy_evict_first();
    const uint64_t pol_slab = (a.variant & 2) ? l2_policy_evict_normal() : l2_policy_evict_last();

    for (int64_t img = blockIdx.x; img < a.n; img += gridDim.x) {
        const uint8_t *src = a.px + (size_t)img * a.img_pitch;
        const uint8_t *next_src = img + gridDim.x < a.n ? src + (size_t)gridDim.x * a.img_pitch : nullptr;
        float esum = 0.0f;   // threads 0..5: running window sum of one edge column, carried from band to band
        for (int b0 = 0; b0 < H; b0 += FBAND) {
            const int rows_out = min(FBAND, H - b0);
            const int Lr0 = b0 - HT;
            const int nL = rows_out + WC - 1;
            int s_begin = 0;
            if (b0 > 0) {
                // the band above already converted this band's first WC - 1 luma rows (its last ones): they and
                // their edge-column entries (plus the row above them) move to the top instead of being read again
                s_begin = WC - 1;
                for (int idx = threadIdx.x; idx < (WC - 1) * (FLP / 16); idx += FTHREADS)
                    reinterpret_cast<uint4 *>(sL)[idx] = reinterpret_cast<const uint4 *>(sL + (size_t)FBAND * FLP)[idx];
                for (int idx = threadIdx.x; idx < 6 * WC; idx += FTHREADS) {
                    const int c = idx / WC, k = idx - c * WC;
                    sE[c * E_PITCH + k] = sE[c * E_PITCH + FBAND + k];
                }
                __syncthreads();
            }
            // (only behind the 2x pre-downsample: a plane read 1:1 has a quarter of the pixels per plane row, its
            // front end is a third of the time and taking turns costs 2.5 % there)
            const bool take_turns = DOWN2 && !(a.variant & 8);
            if (take_turns) {
                // the two CTAs of an SM take turns in the load phase (front end ~ half of a CTA's time): one
                // streams pixels while the other runs its chains, instead of both idling HBM or both queueing
                // on it (+1.5-2 % measured, tools/pdq_variants.py)
                if (threadIdx.x == 0)
                    while (atomicCAS(&g_front_lock[smid], 0u, 1u) != 0u) __nanosleep(100);
                __syncthreads();
            }
            front_end<LAYOUT, DOWN2, PACKED>(src, a.row_pitch, H, Lr0, nL, s_begin, sL, a.pf_mode, a.pf_rows, pol_px);
            __syncthreads();
            if (take_turns && threadIdx.x == 0) atomicExch(&g_front_lock[smid], 0u);
            clk.lap(PH_FRONT);
            edge_inputs(sL, sE, max(s_begin, -Lr0), min(nL, H - Lr0));
            __syncthreads();
            if (threadIdx.x < 6) edge_walk<WC>(sE + threadIdx.x * E_PITCH, H, b0, rows_out, esum);
            __syncthreads();
            clk.lap(PH_EDGE);
            if (warp < NWC) chain_phase<WC>(sL, sE, p3t, H, b0, rows_out, nL, pol_slab);
            // warm L2 with the first PF_ROWS rows of whatever the front end loads next (the next band
            // of this image, else the first band of the CTA's next image), a few us before it starts
            if (lane == 0 && a.pf_mode >= 2) {
                if (b0 + FBAND < H)
                    l2_prefetch_rows<LAYOUT, DOWN2>(src, a.row_pitch, H, b0 + FBAND + HB, b0 + FBAND + HB + a.pf_rows, warp, 8, pol_px);
                else if (next_src != nullptr)
                    l2_prefetch_rows<LAYOUT, DOWN2>(next_src, a.row_pitch, H, 0, a.pf_rows, warp, 8, pol_px);
            }
            __syncthreads();
            clk.lap(PH_CHAIN);
        }
        // pass 4 + decimation into the tail's 64 x 64 buffer, then quality / DCT / hash
        pass4<WC>(p3t, H, ts.B, reinterpret_cast<float *>(smem + P4_STAGE_OFF), ts.T, ts.p4_bar, clk);
        __syncthreads();
        clk.lap(PH_P4_STAGE);   // (the last gather)
        if (!(a.variant & 4)) {
            // the slab has been consumed: drop its (dirty) lines from L2 instead of writing them back to HBM
            const int lines = (H * 4 + 127) / 128;
            for (int idx = threadIdx.x; idx < 64 * lines; idx += FTHREADS) {
                const float *q = p3t + (size_t)(idx / lines) * P3_PITCH + (idx % lines) * 32;
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(q) : "memory");
            }
        }
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

template <int LAYOUT, bool DOWN2, int WC>
int launch_fused(rh_ctx *ctx, const FusedArgs &a, int grid) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    const bool packed = a.row_pitch == (size_t)FW * (DOWN2 ? 2 : 1) * CH;
    auto kern = packed ? pdq_fused_kernel<LAYOUT, DOWN2, WC, true> : pdq_fused_kernel<LAYOUT, DOWN2, WC, false>;
    RH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FSMEM));
    // the per-SM turn locks start free (a launch that died holding one must not block the next)
    void *locks = nullptr;
    RH_CUDA(ctx, cudaGetSymbolAddress(&locks, g_front_lock));
    RH_CUDA(ctx, cudaMemsetAsync(locks, 0, sizeof(g_front_lock), ctx->stream));
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
    void *p, *p_p3t;
    RH_TRY(scratch(ctx, S_W3, (size_t)grid * 64 * P3_PITCH * sizeof(float), &p_p3t));
    FusedArgs a;
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
    a.variant = ctx->pdq_variant;
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
        static const char *names[NPHASE] = {"front", "edge", "chain", "p4_stage", "p4_chain", "tail"};
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
