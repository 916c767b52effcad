// pdq_float.cu -- the second fused PDQ kernel: every plane that pdq_fused.cu does not take (plane
// width != 512: portrait photos, small images; or a 512-wide plane lower than 193 rows), as long as it
// is at most 512 x 512 with a width that is a multiple of 8.  Replaces pdqhash.rs:166-262 for those
// shapes; results are bit-identical to the reference's sequential float arithmetic.
//
// With a row window other than 8 (pdqhash.rs:246: ceil(cols / 64)) pass 1 is no longer exact, so all of
// passes 2, 3 and 4 are genuine sequential float chains (pdqhash.rs:380-387: the rounding error of a
// running sum persists along the line) and are executed as written; what is restructured is where the
// data lives.  The pixels are read ONCE; the generic pipeline this replaces moved ~21 bytes of f32 planes
// per plane pixel through HBM (7 kernels).
//
// Per image (one CTA of 256 threads, 2 CTAs per SM), in bands of 32 luma rows:
//   F    global (128-bit loads, evict-first) -> luma601 -> (2x2 rounded average) -> u8 luma rows in
//        shared memory
//   P12  one thread per two adjacent plane columns, rows in order:
//        pass 1 is a function of the luma row alone: the clipped window sum S is an integer (two DP4A
//        against 0/1 weights, windows of both columns from one 12-byte fetch) and P1 = RN(S / count)
//        by a two-term reciprocal (tests/test_fused_model.py checks every (S, count));
//        pass 2 is the column chain: running sum in a register for the whole image, the values that
//        leave the window come from a ring of the last 8 pass-1 rows in shared memory; the quotient
//        goes to a 32-row f32 band
//   P3   one warp, lane = row of that band, while the other seven warps convert the next band: the row chain
//        across the columns (eight per step from two conflict-free LDS.128, leaving values in registers),
//        the running sums of the 64 decimated columns (pdqhash.rs:439) go to the per-CTA slab in L2 and are
//        divided by their window sizes by the whole CTA inside pass 4 (in shared memory, as each chunk lands)
//   then pass 4 over the slab and the 64x64 -> hash tail, both shared with pdq_fused.cu.
#include "common.cuh"
#include "pdq_pass4.cuh"
#include "tma.cuh"

namespace {

using namespace rh;

constexpr int GBR = 32;               // luma rows per band
constexpr int GLP = 16 + 512 + 16;    // luma row pitch: 16 zero bytes in front of column 0, >= 16 behind the last
constexpr int GP2_ROWS = GBR + 4;     // pass-2 rows a band can produce (the last band adds the shrink-phase rows)

struct FloatArgs {
    const uint8_t *px;
    size_t row_pitch, img_pitch;
    int64_t n;
    int W, H;
    float *p3t;        // [gridDim.x][64][P3_PITCH]
    const float *dct;  // 16 x 64
    TailOut out;
    int64_t out_offset;
    unsigned long long *phase_clk;   // nullptr, or [NPHASE] cycle totals of thread 0 of every CTA ("pdq.phase_clocks")
};

// pitch of the pass-2 band in floats: rows 16-byte aligned and 4 mod 32 apart, so that the row chains
// (lane = row) read four columns with one conflict-free LDS.128
__host__ __device__ inline int p2_pitch(int W) { return ((W + 31) & ~31) + 4; }

__host__ __device__ inline size_t float_smem_bytes(int W) {
    size_t body = (size_t)GBR * GLP + (size_t)8 * W * 4 + (size_t)GP2_ROWS * p2_pitch(W) * 4;
    body = (body + 15) & ~size_t(15);
    if (body < P4_SMEM_BYTES) body = P4_SMEM_BYTES;   // the tail and pass-4 staging alias the band buffers
    return body + 16 * DCT_PITCH * 4;
}

// Phase F for luma rows [i0, i1) of the plane: W / 8 threads per row, 8 plane pixels each, two register
// sets (the loads of the next row are in flight while this one is converted).
// `tid` in [0, nthreads): the first band is converted by the whole CTA, the others by warps 1..7 while warp 0
// runs the previous band's row chains.
template <int LAYOUT, bool DOWN2>
__device__ __noinline__ void front_rows(const uint8_t *__restrict__ src, size_t row_pitch, int W8, int i0, int i1,
                                           uint8_t *sL, uint64_t pol, int tid, int nthreads) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr int BYTES = 8 * SPP * CH;
    constexpr int NW = BYTES / 4;
    const int per_sweep = nthreads / W8;             // rows converted per sweep
    const int rsub = tid / W8, col8 = tid - rsub * W8;
    if (rsub >= per_sweep) return;
    uint32_t w0[2][NW], w1[2][DOWN2 ? NW : 1];
    const size_t rstep = (size_t)per_sweep * SPP * row_pitch;
    int i = i0 + rsub;
    const uint8_t *p = src + (size_t)(i * SPP) * row_pitch + (size_t)col8 * BYTES;
    uint8_t *d = sL + (size_t)(i - i0) * GLP + 16 + col8 * 8;
    if (i < i1) {
        load_px<BYTES>(p, w0[0], pol);
        if (DOWN2) load_px<BYTES>(p + row_pitch, w1[0], pol);
    }
    while (i < i1) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (i + per_sweep * (q + 1) < i1) {
                load_px<BYTES>(p + (q + 1) * rstep, w0[q ^ 1], pol);
                if (DOWN2) load_px<BYTES>(p + (q + 1) * rstep + row_pitch, w1[q ^ 1], pol);
            }
            if (i + per_sweep * q < i1) {
                RH_CHECK_IDX((i - i0) + per_sweep * q, GBR);
                *reinterpret_cast<uint2 *>(d + (size_t)per_sweep * q * GLP) = luma8<LAYOUT, DOWN2, NW>(w0[q], w1[q]);
            }
        }
        i += 2 * per_sweep;
        p += 2 * rstep;
        d += (size_t)2 * per_sweep * GLP;
    }
}

// Per-thread constants of phase P12 for the thread's two columns c0 = 2 t and c0 + 1.
struct ColumnSetup {
    int word, shift;          // the 12 luma bytes that hold both windows: aligned word index within the row, bit shift
    uint32_t m0a, m0b;        // 0/1 byte weights of column c0's window over bytes 0..3 / 4..7 of the fetch
    uint32_t m1a, m1b, m1c;   // same for column c0 + 1 (its window may reach byte 8)
    float y0h, y0l, y1h, y1l; // two-term reciprocals of the two clipped window sizes
};

__device__ __forceinline__ ColumnSetup column_setup(int c0, int W, int wr) {
    const int half = (wr + 2) / 2, ht = wr - half, hb = half - 1;
    ColumnSetup s;
    const int A = 16 + c0 - ht;   // byte offset of the first window byte within the padded luma row
    s.word = A >> 2;
    s.shift = (A & 3) * 8;
    // the fetch starts at column c0 - ht: byte u belongs to column c0's window for u < wr, to column c0 + 1's
    // for 1 <= u <= wr; bytes outside the image are the zero pads of the row
    uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (u < wr) a0 |= 1u << (8 * u);
        if (u + 4 < wr) b0 |= 1u << (8 * u);
        if (u >= 1 && u <= wr) a1 |= 1u << (8 * u);
        if (u + 4 <= wr) b1 |= 1u << (8 * u);
    }
    s.m0a = a0; s.m0b = b0; s.m1a = a1; s.m1b = b1;
    s.m1c = wr == 8 ? 1u : 0u;
    const float d0 = (float)window_count(c0, W, ht, hb), d1 = (float)window_count(min(c0 + 1, W - 1), W, ht, hb);
    s.y0h = __frcp_rn(d0); s.y0l = __fmul_rn(__fmaf_rn(-d0, s.y0h, 1.0f), s.y0h);
    s.y1h = __frcp_rn(d1); s.y1l = __fmul_rn(__fmaf_rn(-d1, s.y1h, 1.0f), s.y1h);
    return s;
}

// pass-1 values of the thread's two columns for one luma row (pdqhash.rs:366-395 along the row: the
// running sums of u8 samples are exact integers, so each output is RN(window sum / window size))
__device__ __forceinline__ float2 pass1_pair(const uint8_t *row, const ColumnSetup &cs) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(row) + cs.word;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    const uint32_t lo = __funnelshift_r(w0, w1, cs.shift), hi = __funnelshift_r(w1, w2, cs.shift);
    const uint32_t ex = (w2 >> cs.shift) & 0xFFu;   // byte 8 of the fetch
    const uint32_t s0 = __dp4a(lo, cs.m0a, __dp4a(hi, cs.m0b, 0u));
    const uint32_t s1 = __dp4a(lo, cs.m1a, __dp4a(hi, cs.m1b, ex * cs.m1c));
    const float f0 = (float)s0, f1 = (float)s1;
    return make_float2(__fmaf_rn(f0, cs.y0h, __fmul_rn(f0, cs.y0l)), __fmaf_rn(f1, cs.y1h, __fmul_rn(f1, cs.y1l)));
}

// Phase P3 for one row of the pass-2 band (one lane): the row pass of box_one_d_float (pdqhash.rs:341-396) over W
// samples with window WR.  A single warp runs this while the other seven convert the next band, so it is written
// for one warp's latency: eight columns per step from two LDS.128, the values that leave the window come from
// registers (WR is a compile-time constant), a window that is still filling subtracts exact zeros (no prologue),
// and the running sums simply replace the samples they were computed from (two STS.128: no per-column test).
// On return q[i] is the window sum after column i entered (= the sum of output column i - HB) and q[W + k] the
// k-th shrink-phase sum; row_samples picks the 64 decimated columns out of that.
template <int WR>
__device__ __noinline__ void row_chain(float *q, int W) {
    constexpr int HALF = (WR + 2) / 2, HB = HALF - 1;
    float p[8];
#pragma unroll
    for (int k = 0; k < 8; k++) p[k] = 0.0f;
    float sum = 0.0f;
    for (int i = 0; i < W; i += 8) {
        const float4 a = *reinterpret_cast<const float4 *>(q + i), b = *reinterpret_cast<const float4 *>(q + i + 4);
        const float n[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float s[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {   // pdqhash.rs:380-387: add the entering sample, then subtract the leaving one
            sum = __fsub_rn(__fadd_rn(sum, n[k]), k >= WR ? n[k - WR] : p[8 + k - WR]);
            s[k] = sum;
        }
        *reinterpret_cast<float4 *>(q + i) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4 *>(q + i + 4) = make_float4(s[4], s[5], s[6], s[7]);
#pragma unroll
        for (int k = 0; k < 8; k++) p[k] = n[k];
    }
#pragma unroll
    for (int k = 0; k < HB; k++) {      // shrink phase (pdqhash.rs:389-395): outputs W-HB .. W-1 (the band's pitch
        sum = __fsub_rn(sum, p[8 - WR + k]);   // leaves >= 4 floats behind column W-1)
        q[W + k] = sum;
    }
}

// the RAW running sums of the 64 decimated columns floor((2 j + 1) W / 128) (pdqhash.rs:439) of one row -> slab;
// they are divided by their window sizes by pass 4, in shared memory, as each staged chunk lands (SlabNorm)
__device__ __forceinline__ void row_samples(const float *q, int W, int hb, float *slab_row, uint64_t pol_slab) {
#pragma unroll 8
    for (int j = 0; j < 64; j++) {
        const int o = ((2 * j + 1) * W) >> 7;
        st_slab(slab_row + (size_t)j * P3_PITCH, q[o + hb], pol_slab);   // o + hb >= W: a shrink-phase sum
    }
}

__device__ __forceinline__ void row_chain_any(int wr, float *q, int W, float *slab_row, uint64_t pol_slab) {
    switch (wr) {
        case 2: row_chain<2>(q, W); break;
        case 3: row_chain<3>(q, W); break;
        case 4: row_chain<4>(q, W); break;
        case 5: row_chain<5>(q, W); break;
        case 6: row_chain<6>(q, W); break;
        case 7: row_chain<7>(q, W); break;
        default: row_chain<8>(q, W); break;
    }
    row_samples(q, W, (wr + 2) / 2 - 1, slab_row, pol_slab);
}

template <int LAYOUT, bool DOWN2, int WC>
__global__ void __launch_bounds__(FTHREADS, 2) pdq_float_kernel(const FloatArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int W = a.W, H = a.H;
    uint8_t *sL = smem;                                                   // [GBR][GLP] luma rows of the band
    float *sP1 = reinterpret_cast<float *>(smem + (size_t)GBR * GLP);     // [8][W]     ring of pass-1 rows
    float *sP2 = sP1 + (size_t)8 * W;                                     // [GP2_ROWS][W + 1] pass-2 band
    TailSmem &ts = *reinterpret_cast<TailSmem *>(smem);                   // aliases all of the above after the last band
    float *sD = reinterpret_cast<float *>(smem + (float_smem_bytes(W) - 16 * DCT_PITCH * 4));   // DCT matrix, resident
    constexpr int HALF = (WC + 2) / 2, HT = WC - HALF, HB = HALF - 1;
    const int wr = (W + 63) >> 6;                                         // pdqhash.rs:246
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W8 = W >> 3, P2P = p2_pitch(W);
    float *p3t = a.p3t + (size_t)blockIdx.x * 64 * P3_PITCH;
    for (int idx = threadIdx.x; idx < 1024; idx += FTHREADS) sD[(idx >> 6) * DCT_PITCH + (idx & 63)] = a.dct[idx];
    PhaseClock clk;
    clk.start(a.phase_clk);
    const uint64_t pol_px = l2_policy_evict_first(), pol_slab = l2_policy_evict_last();
    const int c0 = 2 * (int)threadIdx.x;
    const bool active = c0 < W;
    const ColumnSetup cs = column_setup(active ? c0 : 0, W, wr);

    for (int64_t img = blockIdx.x; img < a.n; img += gridDim.x) {
        const uint8_t *src = a.px + (size_t)img * a.img_pitch;
        // the zero pads of the luma rows (the tail of the previous image aliased them)
        for (int idx = threadIdx.x; idx < GBR * 8; idx += FTHREADS) {
            uint8_t *row = sL + (size_t)(idx >> 3) * GLP;
            const int k = idx & 7;
            if (k < 4)
                reinterpret_cast<uint32_t *>(row)[k] = 0u;
            else
                reinterpret_cast<uint32_t *>(row + 16 + W)[k - 4] = 0u;
        }
        float sum0 = 0.0f, sum1 = 0.0f;   // pass-2 running sums of the thread's two columns
        constexpr bool WC_POW2 = (WC & (WC - 1)) == 0;
        front_rows<LAYOUT, DOWN2>(src, a.row_pitch, W8, 0, min(GBR, H), sL, pol_px, (int)threadIdx.x, FTHREADS);
        __syncthreads();
        clk.lap(PH_FRONT);
        for (int i0 = 0; i0 < H; i0 += GBR) {
            const int i1 = min(i0 + GBR, H);
            const bool last = i1 == H;
            // pull the source rows of the next band (or of the first band of the CTA's next image) into L2
            // while this band's columns are computed: one bulk prefetch per source row
            {
                constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
                constexpr int SPP = DOWN2 ? 2 : 1;
                const uint32_t row_bytes = ((uint32_t)(W * SPP * CH)) & ~15u;
                const uint8_t *base = nullptr;
                int rows = 0;
                if (!last) {
                    base = src + (size_t)(i1 * SPP) * a.row_pitch;
                    rows = (min(i1 + GBR, H) - i1) * SPP;
                } else if (img + gridDim.x < a.n) {
                    base = src + (size_t)gridDim.x * a.img_pitch;
                    rows = min(GBR, H) * SPP;
                }
                if (base != nullptr && (int)threadIdx.x < rows && row_bytes)
                    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(
                                     base + (size_t)threadIdx.x * a.row_pitch),
                                 "r"(row_bytes), "l"(pol_px)
                                 : "memory");
            }
            const int r0 = max(0, i0 - HB);                       // first pass-2 row this band produces
            const int nb = (last ? H : i1 - HB) - r0;             // and how many
            if (active) {
                int i = i0;
                // the first WC rows of the image: the window is still filling (pdqhash.rs:366-378)
                for (; i < min(WC, i1); i++) {
                    const float2 x = pass1_pair(sL + (size_t)(i - i0) * GLP, cs);
                    sum0 = __fadd_rn(sum0, x.x);
                    sum1 = __fadd_rn(sum1, x.y);
                    *reinterpret_cast<float2 *>(sP1 + (size_t)(i & 7) * W + c0) = x;
                    if (i >= HB) {
                        const int o = i - HB, cnt = window_count(o, H, HT, HB);
                        float *dst = sP2 + (size_t)(o - r0) * P2P + c0;
                        RH_CHECK_IDX(o - r0, GP2_ROWS);
                        dst[0] = div_count(sum0, cnt);
                        dst[1] = div_count(sum1, cnt);
                    }
                }
                // steady state (pdqhash.rs:380-387): row i enters, row i - WC leaves, the full window of WC rows is
                // divided.  No branches: the pass-1 work of consecutive rows overlaps, only the add / subtract pairs
                // are sequential.
                const uint8_t *lrow = sL + (size_t)(i - i0) * GLP;
                float *dst = sP2 + (size_t)(i - HB - r0) * P2P + c0;
#pragma unroll 4
                for (; i < i1; i++, lrow += GLP, dst += P2P) {
                    const float2 x = pass1_pair(lrow, cs);
                    const float2 old = *reinterpret_cast<const float2 *>(sP1 + (size_t)((i - WC) & 7) * W + c0);
                    sum0 = __fsub_rn(__fadd_rn(sum0, x.x), old.x);   // add, then subtract
                    sum1 = __fsub_rn(__fadd_rn(sum1, x.y), old.y);
                    *reinterpret_cast<float2 *>(sP1 + (size_t)(i & 7) * W + c0) = x;
                    RH_CHECK_IDX(i - HB - r0, GP2_ROWS);
                    *reinterpret_cast<float2 *>(dst) =
                        make_float2(WC_POW2 ? __fmul_rn(sum0, 1.0f / WC) : __fdiv_rn(sum0, (float)WC),
                                    WC_POW2 ? __fmul_rn(sum1, 1.0f / WC) : __fdiv_rn(sum1, (float)WC));
                }
                if (last) {
                    for (int o = H - HB; o < H; o++) {            // shrink phase (pdqhash.rs:389-395)
                        const float2 old = *reinterpret_cast<const float2 *>(sP1 + (size_t)((o - HT - 1) & 7) * W + c0);
                        sum0 = __fsub_rn(sum0, old.x);
                        sum1 = __fsub_rn(sum1, old.y);
                        const int cnt = window_count(o, H, HT, HB);
                        float *dst = sP2 + (size_t)(o - r0) * P2P + c0;
                        RH_CHECK_IDX(o - r0, GP2_ROWS);
                        dst[0] = div_count(sum0, cnt);
                        dst[1] = div_count(sum1, cnt);
                    }
                }
            }
            __syncthreads();
            clk.lap(PH_EDGE);    // (reported as "p12")
            // warp 0: the row chains of this band; warps 1..7: the luma rows of the next band
            if (warp == 0) {
                for (int base = 0; base < nb; base += 32) {
                    const int lr = base + lane;
                    if (lr < nb) row_chain_any(wr, sP2 + (size_t)lr * P2P, W, p3t + r0 + lr, pol_slab);
                }
                clk.lap(PH_CHAIN);   // warp 0's own time in the row chains
            } else if (!last) {
                long long t0 = 0;
                if (a.phase_clk != nullptr && threadIdx.x == 32) t0 = clock64();
                front_rows<LAYOUT, DOWN2>(src, a.row_pitch, W8, i1, min(i1 + GBR, H), sL, pol_px, (int)threadIdx.x - 32,
                                          FTHREADS - 32);
                if (a.phase_clk != nullptr && threadIdx.x == 32)
                    atomicAdd(a.phase_clk + PH_P4_CHAIN, (unsigned long long)(clock64() - t0));   // warp 1's time in the next band's front end
            }
            __syncthreads();
            clk.lap(PH_P4_STAGE);   // warp 0 waiting for the front end (or vice versa)
        }
        // pass 4 (which first turns the slab's raw pass-3 sums into quotients, chunk by chunk in shared memory) +
        // decimation into the tail's 64 x 64 buffer, then quality / DCT / hash
        pass4<WC>(p3t, H, ts.B, reinterpret_cast<float *>(smem + P4_STAGE_OFF), ts.T, ts.p4_bar, clk, SlabNorm{W, wr});
        __syncthreads();
        {
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

template <int LAYOUT, bool DOWN2>
int launch_float(rh_ctx *ctx, const FloatArgs &a, int grid, int wc) {
    const size_t smem = float_smem_bytes(a.W);
#define RH_FLOAT_CASE(WCV)                                                                                     \
    case WCV: {                                                                                                \
        auto kern = pdq_float_kernel<LAYOUT, DOWN2, WCV>;                                                      \
        RH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<grid, FTHREADS, smem, ctx->stream>>>(a);                                                        \
        break;                                                                                                 \
    }
    switch (wc) {
        RH_FLOAT_CASE(2)
        RH_FLOAT_CASE(3)
        RH_FLOAT_CASE(4)
        RH_FLOAT_CASE(5)
        RH_FLOAT_CASE(6)
        RH_FLOAT_CASE(7)
        RH_FLOAT_CASE(8)
        default:
            return fail(ctx, RH_EUNSUPPORTED, "float PDQ kernel: column window not instantiated");
    }
#undef RH_FLOAT_CASE
    RH_LAUNCHED(ctx, "pdq_float_kernel");
    return RH_OK;
}

}  // namespace

namespace rh {

// planes of 72..512 columns (a multiple of 8) and 65..512 rows: row and column windows 2..8
int pdq_float_supported(int W, int H) { return W >= 72 && W <= 512 && (W & 7) == 0 && H >= 65 && H <= 512; }

int pdq_float_run(rh_ctx *ctx, const uint8_t *d_px, int layout, bool down2, int64_t n, int W, int H, size_t row_pitch,
                  size_t img_pitch, const TailOut &out, int64_t out_offset, const float *d_dct) {
    if (!pdq_float_supported(W, H)) return fail(ctx, RH_EUNSUPPORTED, "float PDQ kernel: unsupported plane size");
    if ((reinterpret_cast<uintptr_t>(d_px) | row_pitch | img_pitch) & 15)
        return fail(ctx, RH_EINVAL, "float PDQ kernel: pixels must be 16-byte aligned");
    int grid = ctx->sm_count * 2;
    if (grid > n) grid = (int)n;
    void *p_p3t;
    RH_TRY(scratch(ctx, S_W3, (size_t)grid * 64 * P3_PITCH * sizeof(float), &p_p3t));
    FloatArgs a;
    a.px = d_px;
    a.row_pitch = row_pitch;
    a.img_pitch = img_pitch;
    a.n = n;
    a.W = W;
    a.H = H;
    a.p3t = (float *)p_p3t;
    a.dct = d_dct;
    a.out = out;
    a.out_offset = out_offset;
    a.phase_clk = nullptr;
    if (ctx->pdq_phase_clocks) {
        void *p;
        RH_TRY(scratch(ctx, S_W8, NPHASE * sizeof(unsigned long long), &p));
        a.phase_clk = (unsigned long long *)p;
        RH_CUDA(ctx, cudaMemsetAsync(p, 0, NPHASE * sizeof(unsigned long long), ctx->stream));
    }
    const int wc = (H + 63) / 64;
    int rc;
    if (layout == RH_LAYOUT_RGB8)
        rc = down2 ? launch_float<RH_LAYOUT_RGB8, true>(ctx, a, grid, wc) : launch_float<RH_LAYOUT_RGB8, false>(ctx, a, grid, wc);
    else if (layout == RH_LAYOUT_RGBA8)
        rc = down2 ? launch_float<RH_LAYOUT_RGBA8, true>(ctx, a, grid, wc) : launch_float<RH_LAYOUT_RGBA8, false>(ctx, a, grid, wc);
    else
        rc = down2 ? launch_float<RH_LAYOUT_LUMA8, true>(ctx, a, grid, wc) : launch_float<RH_LAYOUT_LUMA8, false>(ctx, a, grid, wc);
    if (rc == RH_OK && a.phase_clk) {
        unsigned long long h[NPHASE];
        RH_CUDA(ctx, cudaMemcpyAsync(h, a.phase_clk, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        static const char *names[NPHASE] = {"front0", "p12", "p3(warp0)", "wait+pass4", "front(warp1)", "tail"};
        unsigned long long tot = 0;
        for (int i = 0; i < NPHASE; i++) tot += h[i];
        fprintf(stderr, "[pdq_float phases] n=%lld %dx%d", (long long)n, W, H);
        for (int i = 0; i < NPHASE; i++)
            fprintf(stderr, "  %s %.0f cyc/img (%.1f%%)", names[i], (double)h[i] / (double)n, 100.0 * (double)h[i] / (double)tot);
        fprintf(stderr, "\n");
    }
    return rc;
}

}  // namespace rh
