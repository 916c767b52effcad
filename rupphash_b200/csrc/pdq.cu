// pdq.cu -- hot path #1: batched PDQ hashing (pdqhash.rs).
//
// Entry points: rh_pdq_hash_batch, rh_pdq_from_buffer64, rh_pdq_hash_from_coeffs,
// rh_pdq_dihedral_from_coeffs.  Two device pipelines produce the 64 x 64 buffer:
//   * the fused kernel of pdq_fused.cu for planes 512 px wide and 193..512 px high (both BASELINE
//     shapes; 8:3 to 1:1 landscape shapes in general),
//   * the generic pipeline below for every other supported size: one thread walks one line
//     exactly like box_one_d_float (pdqhash.rs:341-396), planes live in device scratch.
// Both end in pdq_tail.cuh.  Arithmetic is f32 with explicit round-to-nearest mul/add/div
// intrinsics, so nvcc can never contract a*b+c into an FMA (Rust never does).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "pdq_luma.cuh"
#include "pdq_tail.cuh"

namespace rh {
int pdq_fused_supported(int W, int H);
int pdq_fused_aligned(const void *px, size_t row_pitch, size_t img_pitch);
int pdq_float_supported(int W, int H);
int pdq_float_run(rh_ctx *ctx, const uint8_t *d_px, int layout, bool down2, int64_t n, int W, int H, size_t row_pitch,
                  size_t img_pitch, const TailOut &out, int64_t out_offset, const float *d_dct);
int pdq_fused_run(rh_ctx *ctx, const uint8_t *d_px, int layout, bool down2, int64_t n, int W, int H, size_t row_pitch,
                  size_t img_pitch, const TailOut &out, int64_t out_offset, const float *d_dct);
}  // namespace rh

namespace {

using namespace rh;

// pdqhash.rs:287-304 -- computed on the host with the same f32 steps and the host libm cosf
// (Rust's f32::cos lowers to the platform cosf), rows are frequencies 1..16.
void host_dct_matrix(float *D) {
    const float PI_F = 3.14159265358979323846f;
    const float num_cols = 64.0f;
    const float inv_sqrt_cols = 1.0f / sqrtf(num_cols);
    const float sqrt_2 = sqrtf(2.0f);
    for (int i = 0; i < 16; i++) {
        const float freq = (float)(i + 1);
        const float norm = inv_sqrt_cols * sqrt_2;
        for (int j = 0; j < 64; j++) {
            volatile float a = PI_F * freq;
            volatile float b = 2.0f * (float)j + 1.0f;
            volatile float num = a * b;
            volatile float angle = num / (2.0f * num_cols);
            D[i * 64 + j] = norm * cosf(angle);
        }
    }
}

// pdqhash.rs:268-284: (299 r + 587 g + 114 b + 500) / 1000 with truncating u32 division.
// floor(v / 1000) == umulhi(v, ceil(2^32 / 1000)) for every v < 6.1e6 (v <= 255 500 here).
__device__ __forceinline__ uint32_t luma601(uint32_t r, uint32_t g, uint32_t b) {
    return __umulhi(299u * r + 587u * g + 114u * b + 500u, 4294968u);
}

template <int LAYOUT>
__device__ __forceinline__ uint32_t load_luma(const uint8_t *p) {
    if (LAYOUT == RH_LAYOUT_LUMA8) return p[0];
    return luma601(p[0], p[1], p[2]);
}

// Generic front end: one thread per output pixel of the (optionally 2x reduced) luma plane.
// DOWN2 is fast_image_resize's Box convolution at an exact 2x ratio: two taps of 1/2 per axis,
// horizontal pass first, each pass rounding half up to u8 (pdqhash.rs:203-220).
template <int LAYOUT, bool DOWN2>
__global__ void luma_plane_kernel(const uint8_t *__restrict__ px, size_t row_pitch, size_t img_pitch, int n, int W,
                                  int H, uint8_t *__restrict__ L) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t per = (size_t)W * H;
    if (idx >= per * n) return;
    const size_t img = idx / per;
    const int y = (int)((idx % per) / W), x = (int)(idx % W);
    const uint8_t *base = px + img * img_pitch;
    uint32_t v;
    if (DOWN2) {
        const uint8_t *r0 = base + (size_t)(2 * y) * row_pitch + (size_t)(2 * x) * CH;
        const uint8_t *r1 = r0 + row_pitch;
        const uint32_t h0 = (load_luma<LAYOUT>(r0) + load_luma<LAYOUT>(r0 + CH) + 1u) >> 1;
        const uint32_t h1 = (load_luma<LAYOUT>(r1) + load_luma<LAYOUT>(r1 + CH) + 1u) >> 1;
        v = (h0 + h1 + 1u) >> 1;
    } else {
        v = load_luma<LAYOUT>(base + (size_t)y * row_pitch + (size_t)x * CH);
    }
    L[idx] = (uint8_t)v;
}

// The same front end, eight plane pixels per thread from 128-bit loads (pdq_luma.cuh: the fused kernel's
// DP2A / FMA luma): for 16-byte aligned rows and plane widths that are multiples of 8.
template <int LAYOUT, bool DOWN2>
__global__ void luma_plane8_kernel(const uint8_t *__restrict__ px, size_t row_pitch, size_t img_pitch, int n, int W,
                                   int H, uint8_t *__restrict__ L) {
    constexpr int CH = LAYOUT == RH_LAYOUT_RGB8 ? 3 : (LAYOUT == RH_LAYOUT_RGBA8 ? 4 : 1);
    constexpr int SPP = DOWN2 ? 2 : 1;
    constexpr int BYTES = 8 * SPP * CH;
    constexpr int NW = BYTES / 4;
    const int w8 = W / 8;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t per = (size_t)w8 * H;
    if (idx >= per * n) return;
    const size_t img = idx / per;
    const int y = (int)((idx % per) / w8), x8 = (int)(idx % w8);
    const uint8_t *p = px + img * img_pitch + (size_t)(y * SPP) * row_pitch + (size_t)x8 * BYTES;
    uint32_t w0[NW], w1[DOWN2 ? NW : 1];
    load_chunk<BYTES>(p, w0);
    if (DOWN2) load_chunk<BYTES>(p + row_pitch, w1);
    *reinterpret_cast<uint2 *>(L + img * (size_t)W * H + (size_t)y * W + 8 * x8) = luma8<LAYOUT, DOWN2, NW>(w0, w1);
}

// u8 plane [R][C] -> [C][R] per image: 64 x 64 tiles through shared memory, 32-bit accesses on both
// sides when R and C are multiples of 4 (byte accesses otherwise).
__global__ void __launch_bounds__(256) transpose_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int R, int C) {
    __shared__ uint8_t tile[64][64 + 4];
    const size_t img = blockIdx.z;
    const uint8_t *src = in + img * (size_t)R * C;
    uint8_t *dst = out + img * (size_t)R * C;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 bytes each per step
    const bool vec = ((R | C) & 3) == 0;
    for (int k = ty; k < 64; k += 16) {
        const int r = r0 + k, c = c0 + 4 * tx;
        if (r >= R) continue;
        if (vec && c + 3 < C) {
            *reinterpret_cast<uint32_t *>(&tile[k][4 * tx]) = *reinterpret_cast<const uint32_t *>(src + (size_t)r * C + c);
        } else {
            for (int q = 0; q < 4; q++)
                if (c + q < C) tile[k][4 * tx + q] = src[(size_t)r * C + c + q];
        }
    }
    __syncthreads();
    for (int k = ty; k < 64; k += 16) {
        const int c = c0 + k, r = r0 + 4 * tx;   // output row c, output columns r .. r + 3
        if (c >= C) continue;
        if (vec && r + 3 < R) {
            const uint32_t v = (uint32_t)tile[4 * tx][k] | ((uint32_t)tile[4 * tx + 1][k] << 8) |
                               ((uint32_t)tile[4 * tx + 2][k] << 16) | ((uint32_t)tile[4 * tx + 3][k] << 24);
            *reinterpret_cast<uint32_t *>(dst + (size_t)c * R + r) = v;
        } else {
            for (int q = 0; q < 4; q++)
                if (r + q < R) dst[(size_t)c * R + r + q] = tile[4 * tx + q][k];
        }
    }
}

// One pass of the Jarosz filter: box_one_d_float (pdqhash.rs:341-396) verbatim -- sequential running
// sum, IEEE division by the current window, four phases -- along the ROW index i of a row-major
// [R][C] plane.  One thread per column j, so every load is coalesced across the warp.  Output modes:
//   WALK_T      the whole result, transposed ([C][R]): a warp stages 32 x 32 outputs in shared memory and
//               writes them as 32 coalesced rows, which turns the next pass (along the other axis) into
//               the same coalesced walk;
//   WALK_DEC_T  only the 64 decimation samples of each line (index ((2k+1) R) / 128, pdqhash.rs:435-440),
//               staged and written as [C][64] -- pass 4 only ever reads those columns of pass 3;
//   WALK_DEC    only the 64 decimation samples, written as [64][C] -- the 64 x 64 buffer of the tail.
// In the decimated modes the division is done for the kept samples only.
// generic_chunk chains four walks: L^T -> A -> B^T -> A3 [H][64] -> B64 [64][64].
constexpr int WALK_WARPS = 4;
enum { WALK_T = 0, WALK_DEC_T = 1, WALK_DEC = 2 };
template <typename Tin, int MODE>
__global__ void __launch_bounds__(32 * WALK_WARPS) box_walk_kernel(const Tin *__restrict__ in, float *__restrict__ out, int n,
                                                                 int R, int C, int win) {
    constexpr int TPITCH = MODE == WALK_DEC_T ? 65 : 33;
    __shared__ float tiles[WALK_WARPS][32 * TPITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ctiles = (C + 31) / 32;
    const size_t wid = (size_t)blockIdx.x * WALK_WARPS + warp;   // (image, column tile)
    if (wid >= (size_t)n * ctiles) return;                       // whole warps leave: no block-wide barrier below
    const size_t img = wid / ctiles;
    const int j0 = (int)(wid % ctiles) * 32;
    const int j = min(j0 + lane, C - 1);                         // lanes past the edge shadow the last column
    const bool valid = j0 + lane < C;
    const Tin *src = in + img * (size_t)R * C + j;
    float *dst = out + img * (MODE == WALK_T ? (size_t)R * C : (size_t)64 * C);
    float *tile = tiles[warp];
    int next_k = 0, next_idx = R >> 7;   // decimated modes: the next sample to keep and its line index
    const int lim = R > 1 ? R : 1;
    win = win < 1 ? 1 : (win > lim ? lim : win);
    const int half = (win + 2) / 2;
    const int phase_1 = half - 1, phase_2 = win - half + 1, phase_3 = R > win ? R - win : 0, phase_4 = half - 1;
    int li = 0, ri = 0, oi = 0;
    float sum = 0.0f, curr = 0.0f;
    // one output of row oi = sum / curr
    auto emit = [&](float s, float c) {
        if (MODE == WALK_T) {   // staged, flushed when the 32-row tile (or the plane) is complete
            tile[(oi & 31) * TPITCH + lane] = __fdiv_rn(s, c);
            if ((oi & 31) == 31 || oi == R - 1) {
                const int o0 = oi & ~31, cnt = (oi & 31) + 1;
                __syncwarp();
                for (int k = 0; k < 32; k++) {
                    const int jj = j0 + k;
                    if (jj < C && lane < cnt) dst[(size_t)jj * R + o0 + lane] = tile[lane * TPITCH + k];
                }
                __syncwarp();
            }
        } else if (oi == next_idx) {   // (uniform across the warp)
            const float v = __fdiv_rn(s, c);
            while (next_k < 64 && next_idx == oi) {   // short lines map several samples to one index
                if (MODE == WALK_DEC_T)
                    tile[lane * TPITCH + next_k] = v;
                else if (valid)
                    dst[(size_t)next_k * C + j] = v;
                next_k++;
                next_idx = ((2 * next_k + 1) * R) >> 7;
            }
        }
        oi++;
    };
    for (int i = 0; i < phase_1; i++) {
        sum = __fadd_rn(sum, (float)src[(size_t)ri * C]);
        curr += 1.0f;
        ri++;
    }
    for (int i = 0; i < phase_2; i++) {
        sum = __fadd_rn(sum, (float)src[(size_t)ri * C]);
        curr += 1.0f;
        emit(sum, curr);
        ri++;
    }
    int i3 = 0;
    if (phase_3 >= 8) {   // eight steps at a time, the sixteen loads of the NEXT eight already in flight
        float xin[8], xout[8], nin[8], nout[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            xin[k] = (float)src[(size_t)(ri + k) * C];
            xout[k] = (float)src[(size_t)(li + k) * C];
        }
        for (; i3 + 8 <= phase_3; i3 += 8) {
            const bool more = i3 + 16 <= phase_3;
            if (more) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    nin[k] = (float)src[(size_t)(ri + 8 + k) * C];
                    nout[k] = (float)src[(size_t)(li + 8 + k) * C];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                sum = __fadd_rn(sum, xin[k]);
                sum = __fsub_rn(sum, xout[k]);
                emit(sum, curr);
            }
            li += 8;
            ri += 8;
            if (more) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    xin[k] = nin[k];
                    xout[k] = nout[k];
                }
            }
        }
    }
    for (; i3 < phase_3; i3++) {
        sum = __fadd_rn(sum, (float)src[(size_t)ri * C]);
        sum = __fsub_rn(sum, (float)src[(size_t)li * C]);
        emit(sum, curr);
        li++;
        ri++;
    }
    for (int i = 0; i < phase_4; i++) {
        sum = __fsub_rn(sum, (float)src[(size_t)li * C]);
        curr -= 1.0f;
        emit(sum, curr);
        li++;
    }
    if (MODE == WALK_DEC_T) {   // the 64 samples of each of the warp's 32 lines, one 256-byte row per line
        __syncwarp();
        for (int l = 0; l < 32; l++) {
            if (j0 + l >= C) break;
            dst[(size_t)(j0 + l) * 64 + lane] = tile[l * TPITCH + lane];
            dst[(size_t)(j0 + l) * 64 + 32 + lane] = tile[l * TPITCH + 32 + lane];
        }
    }
}

enum TailSrc { SRC_PLANE = 0, SRC_BUF64 = 1, SRC_COEFFS = 2 };

// One CTA per image.  SRC_PLANE decimates the filtered plane (pdqhash.rs:428-443) on load.
template <int SRC>
__global__ void __launch_bounds__(TAIL_THREADS) pdq_tail_kernel(const float *__restrict__ src, int W, int H,
                                                                const float *__restrict__ dct, TailOut o,
                                                                int64_t out_offset) {
    __shared__ TailSmem s;
    const size_t img = blockIdx.x;
    const size_t oimg = img + (size_t)out_offset;
    if (SRC == SRC_COEFFS) {
        s.C[threadIdx.x] = src[img * 256 + threadIdx.x];
        __syncthreads();
        tail_hashes(s, o, oimg);
        return;
    }
    for (int idx = threadIdx.x; idx < 4096; idx += TAIL_THREADS) {
        if (SRC == SRC_PLANE) {
            const int i = idx >> 6, j = idx & 63;
            const int ini = ((2 * i + 1) * H) / 128, inj = ((2 * j + 1) * W) / 128;
            s.B[idx] = src[img * (size_t)W * H + (size_t)ini * W + inj];
        } else {
            s.B[idx] = src[img * 4096 + idx];
        }
    }
    for (int idx = threadIdx.x; idx < 1024; idx += TAIL_THREADS) s.D[(idx >> 6) * DCT_PITCH + (idx & 63)] = dct[idx];
    __syncthreads();
    const float q = tail_quality(s);
    if (threadIdx.x == 0 && o.quality) o.quality[oimg] = q;
    tail_dct(s, s.D);
    if (o.coeffs) o.coeffs[oimg * 256 + threadIdx.x] = s.C[threadIdx.x];
    tail_hashes(s, o, oimg);
}

// ---- general Box pre-downsample (pdqhash.rs:203-220 -> fast_image_resize 6.1.0, U8 Convolution(Box)) ----
// The crate is not in the reference tree; this follows the oracle's restatement of its
// Pillow-derived algorithm (oracle_pdq.c orc_resize_box_u8): per axis, f64 box weights normalised
// per output pixel, quantised to i16 at the largest precision that keeps the biggest coefficient
// below 2^15; horizontal pass first into a u8 plane, then vertical, each pass computing
// (sum(px * k) + (1 << (p-1))) >> p clipped to [0, 255].  PARITY UNPINNED w.r.t. the real crate.
struct BoxCoeffs {
    std::vector<int> xmin, cnt;
    std::vector<int16_t> k;
    int ksize = 0, precision = 0;
};

void box_precompute(int in_size, int out_size, BoxCoeffs *bc) {
    const double scale = (double)in_size / (double)out_size;
    const double fscale = scale < 1.0 ? 1.0 : scale;
    const double support = 0.5 * fscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    std::vector<double> pre((size_t)out_size * ksize, 0.0);
    bc->xmin.assign(out_size, 0);
    bc->cnt.assign(out_size, 0);
    bc->k.assign((size_t)out_size * ksize, 0);
    bc->ksize = ksize;
    const double ss = 1.0 / fscale;
    double maxk = 0.0;
    for (int o = 0; o < out_size; o++) {
        const double center = (o + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        const int cnt = xmax - xmin;
        double ww = 0.0;
        double *kk = pre.data() + (size_t)o * ksize;
        for (int x = 0; x < cnt; x++) {
            const double t = (x + xmin - center + 0.5) * ss;
            const double wgt = (t > -0.5 && t <= 0.5) ? 1.0 : 0.0;
            kk[x] = wgt;
            ww += wgt;
        }
        for (int x = 0; x < cnt; x++) {
            if (ww != 0.0) kk[x] /= ww;
            if (kk[x] > maxk) maxk = kk[x];
        }
        bc->xmin[o] = xmin;
        bc->cnt[o] = cnt;
    }
    int p;
    for (p = 0; p < 32 - 8 - 2; p++) {
        const int next = (int)(0.5 + maxk * (double)(1 << (p + 1)));
        if (next >= (1 << 15)) break;
    }
    bc->precision = p;
    for (size_t i = 0; i < pre.size(); i++) {
        const double v = pre[i] * (double)(1 << p);
        bc->k[i] = (int16_t)(v < 0 ? (int)(v - 0.5) : (int)(v + 0.5));
    }
}

// one thread per output sample; `line_stride` / `tap_stride` walk the source along the filtered axis
__global__ void box_resize_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n_img,
                                  int out_w, int out_h, size_t src_img, int src_row, int tap_stride, bool vertical,
                                  int dst_row, size_t dst_img,
                                  const int *__restrict__ xmin, const int *__restrict__ cnt,
                                  const int16_t *__restrict__ k, int ksize, int precision) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t per = (size_t)out_w * out_h;
    if (idx >= per * n_img) return;
    const size_t img = idx / per;
    const int y = (int)((idx % per) / out_w), x = (int)(idx % out_w);
    const int o = vertical ? y : x;
    const uint8_t *p = src + img * src_img + (vertical ? (size_t)xmin[o] * src_row + x : (size_t)y * src_row + xmin[o]);
    const int16_t *kk = k + (size_t)o * ksize;
    int acc = 1 << (precision - 1);
    for (int t = 0; t < cnt[o]; t++) acc += (int)p[(size_t)t * tap_stride] * (int)kk[t];
    acc >>= precision;
    dst[img * dst_img + (size_t)y * dst_row + x] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}

struct DeviceBox {
    const int *xmin, *cnt;
    const int16_t *k;
    int ksize, precision;
};

int upload_box(rh_ctx *ctx, const BoxCoeffs &bc, int slot, DeviceBox *out) {
    const size_t n = bc.xmin.size();
    const size_t bytes = n * 8 + bc.k.size() * 2;
    void *p;
    RH_TRY(scratch(ctx, slot, bytes, &p));
    std::vector<uint8_t> host(bytes);
    memcpy(host.data(), bc.xmin.data(), n * 4);
    memcpy(host.data() + n * 4, bc.cnt.data(), n * 4);
    memcpy(host.data() + n * 8, bc.k.data(), bc.k.size() * 2);
    RH_CUDA(ctx, cudaMemcpyAsync(p, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `host` is a stack-lifetime object
    out->xmin = (const int *)p;
    out->cnt = out->xmin + n;
    out->k = (const int16_t *)((const uint8_t *)p + n * 8);
    out->ksize = bc.ksize;
    out->precision = bc.precision;
    return RH_OK;
}

int ensure_dct(rh_ctx *ctx, const float **d_dct) {
    void *p;
    RH_TRY(scratch(ctx, S_DCT, 1024 * sizeof(float), &p));
    if (!ctx->dct_ready) {
        float D[1024];
        host_dct_matrix(D);
        RH_CUDA(ctx, cudaMemcpyAsync(p, D, sizeof D, cudaMemcpyHostToDevice, ctx->stream));
        RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dct_ready = true;
    }
    *d_dct = (const float *)p;
    return RH_OK;
}

// pdqhash.rs:224-235
void target_dimensions(int w, int h, int max_dim, int *nw, int *nh) {
    if (w > h) {
        long long v = (long long)h * max_dim / w;
        *nw = max_dim;
        *nh = (int)(v > 1 ? v : 1);
    } else {
        long long v = (long long)w * max_dim / h;
        *nw = (int)(v > 1 ? v : 1);
        *nh = max_dim;
    }
}

template <int LAYOUT>
int launch_luma(rh_ctx *ctx, bool down2, const uint8_t *px, size_t row_pitch, size_t img_pitch, int n, int W, int H,
                uint8_t *L) {
    const size_t total = (size_t)n * W * H;
    if ((W & 7) == 0 && ((reinterpret_cast<uintptr_t>(px) | row_pitch | img_pitch) & 15) == 0) {
        if (down2)
            luma_plane8_kernel<LAYOUT, true><<<cdiv(total / 8, 256), 256, 0, ctx->stream>>>(px, row_pitch, img_pitch, n, W, H, L);
        else
            luma_plane8_kernel<LAYOUT, false><<<cdiv(total / 8, 256), 256, 0, ctx->stream>>>(px, row_pitch, img_pitch, n, W, H, L);
        RH_LAUNCHED(ctx, "luma_plane8_kernel");
        return RH_OK;
    }
    if (down2)
        luma_plane_kernel<LAYOUT, true><<<cdiv(total, 256), 256, 0, ctx->stream>>>(px, row_pitch, img_pitch, n, W, H, L);
    else
        luma_plane_kernel<LAYOUT, false><<<cdiv(total, 256), 256, 0, ctx->stream>>>(px, row_pitch, img_pitch, n, W, H, L);
    RH_LAUNCHED(ctx, "luma_plane_kernel");
    return RH_OK;
}

// generic pipeline over one chunk of device-resident images
int generic_chunk(rh_ctx *ctx, const uint8_t *d_px, int layout, bool down2, int n, int W, int H, size_t row_pitch,
                  size_t img_pitch, const TailOut &out, int64_t out_offset, const float *d_dct) {
    cudaStream_t st = ctx->stream;
    const size_t plane = (size_t)W * H;
    void *p;
    RH_TRY(scratch(ctx, S_W0, plane * n, &p));
    uint8_t *L = (uint8_t *)p;
    // A also takes pass 3's decimated output ([H][64]), B the final 64 x 64 buffers
    const size_t a_floats = plane > (size_t)64 * H ? plane : (size_t)64 * H, b_floats = plane > 4096 ? plane : 4096;
    RH_TRY(scratch(ctx, S_W1, a_floats * n * sizeof(float), &p));
    float *A = (float *)p;
    RH_TRY(scratch(ctx, S_W2, b_floats * n * sizeof(float), &p));
    float *B = (float *)p;
    if (layout == RH_LAYOUT_RGB8)
        RH_TRY(launch_luma<RH_LAYOUT_RGB8>(ctx, down2, d_px, row_pitch, img_pitch, n, W, H, L));
    else if (layout == RH_LAYOUT_RGBA8)
        RH_TRY(launch_luma<RH_LAYOUT_RGBA8>(ctx, down2, d_px, row_pitch, img_pitch, n, W, H, L));
    else
        RH_TRY(launch_luma<RH_LAYOUT_LUMA8>(ctx, down2, d_px, row_pitch, img_pitch, n, W, H, L));
    const int w_rows = (W + 63) / 64, w_cols = (H + 63) / 64;  // pdqhash.rs:246-247
    // rep 1 (pdqhash.rs:422-425): rows L -> A, cols A -> B; rep 2: rows B -> A, cols A -> B.  Every pass
    // is the coalesced walk along the row index, so the row passes run on transposed planes:
    //   L [H][W] -> L^T [W][H] -(rows, win w_rows)-> A [H][W] -(cols, w_cols)-> B^T [W][H]
    //     -(rows, the 64 decimated columns only)-> A3 [H][64] -(cols, the 64 decimated rows only)-> B64 [64][64]
    RH_TRY(scratch(ctx, S_W9, plane * n, &p));
    uint8_t *LT = (uint8_t *)p;
    transpose_u8_kernel<<<dim3(cdiv(W, 64), cdiv(H, 64), n), 256, 0, st>>>(L, LT, H, W);
    RH_LAUNCHED(ctx, "transpose_u8_kernel");
    const unsigned g_rows = cdiv((size_t)n * cdiv(H, 32), WALK_WARPS), g_cols = cdiv((size_t)n * cdiv(W, 32), WALK_WARPS);
    box_walk_kernel<uint8_t, WALK_T><<<g_rows, 32 * WALK_WARPS, 0, st>>>(LT, A, n, W, H, w_rows);
    RH_LAUNCHED(ctx, "box_walk_kernel");
    box_walk_kernel<float, WALK_T><<<g_cols, 32 * WALK_WARPS, 0, st>>>(A, B, n, H, W, w_cols);
    RH_LAUNCHED(ctx, "box_walk_kernel");
    box_walk_kernel<float, WALK_DEC_T><<<g_rows, 32 * WALK_WARPS, 0, st>>>(B, A, n, W, H, w_rows);
    RH_LAUNCHED(ctx, "box_walk_kernel");
    box_walk_kernel<float, WALK_DEC><<<cdiv((size_t)n * 2, WALK_WARPS), 32 * WALK_WARPS, 0, st>>>(A, B, n, H, 64, w_cols);
    RH_LAUNCHED(ctx, "box_walk_kernel");
    pdq_tail_kernel<SRC_BUF64><<<n, TAIL_THREADS, 0, st>>>(B, W, H, d_dct, out, out_offset);
    RH_LAUNCHED(ctx, "pdq_tail_kernel");
    return RH_OK;
}

__global__ void fill_u8_kernel(uint8_t *p, size_t n, uint8_t v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// enqueue-only core of rh_pdq_hash_batch; `wait` = synchronise and read the kernel time before returning
static int pdq_hash_batch_impl(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h, size_t row_pitch,
                               size_t img_pitch, uint8_t *out_hash, float *out_quality, float *out_coeffs,
                               uint8_t *out_dihedral, uint8_t *out_valid, bool wait) {
    if (!ctx) return RH_EINVAL;
    if (n < 0 || w <= 0 || h <= 0 || (n > 0 && !pixels)) return fail(ctx, RH_EINVAL, "rh_pdq_hash_batch: bad arguments");
    if (layout != RH_LAYOUT_RGB8 && layout != RH_LAYOUT_RGBA8 && layout != RH_LAYOUT_LUMA8)
        return fail(ctx, RH_EINVAL, "rh_pdq_hash_batch: unknown layout");
    const int ch = layout == RH_LAYOUT_RGB8 ? 3 : (layout == RH_LAYOUT_RGBA8 ? 4 : 1);
    if (row_pitch == 0) row_pitch = (size_t)w * ch;
    if (img_pitch == 0) img_pitch = row_pitch * h;
    if (row_pitch < (size_t)w * ch || img_pitch < row_pitch * (size_t)(h - 1) + (size_t)w * ch)
        return fail(ctx, RH_EINVAL, "rh_pdq_hash_batch: pitch smaller than the image");
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->last_ms = 0.0;
    ctx->last_units = 0.0;
    if (n == 0) return RH_OK;

    OutBuf<uint8_t> o_hash, o_dih, o_valid;
    OutBuf<float> o_q, o_c;
    RH_TRY(o_hash.prepare(ctx, out_hash, (size_t)n * 32, S_OUT0));
    RH_TRY(o_q.prepare(ctx, out_quality, (size_t)n, S_OUT1));
    RH_TRY(o_c.prepare(ctx, out_coeffs, (size_t)n * 256, S_OUT2));
    RH_TRY(o_dih.prepare(ctx, out_dihedral, (size_t)n * 256, S_OUT3));
    RH_TRY(o_valid.prepare(ctx, out_valid, (size_t)n, S_OUT4));

    if (w < 5 || h < 5) {  // pdqhash.rs:167-169 -> None for every image of the batch
        if (o_valid.dev) RH_CUDA(ctx, cudaMemsetAsync(o_valid.dev, 0, (size_t)n, st));
        if (o_hash.dev) RH_CUDA(ctx, cudaMemsetAsync(o_hash.dev, 0, (size_t)n * 32, st));
        if (o_q.dev) RH_CUDA(ctx, cudaMemsetAsync(o_q.dev, 0, (size_t)n * 4, st));
        if (o_c.dev) RH_CUDA(ctx, cudaMemsetAsync(o_c.dev, 0, (size_t)n * 1024, st));
        if (o_dih.dev) RH_CUDA(ctx, cudaMemsetAsync(o_dih.dev, 0, (size_t)n * 256, st));
        RH_TRY(o_valid.finish(ctx));
        RH_TRY(o_hash.finish(ctx));
        RH_TRY(o_q.finish(ctx));
        RH_TRY(o_c.finish(ctx));
        RH_TRY(o_dih.finish(ctx));
        if (wait) RH_CUDA(ctx, cudaStreamSynchronize(st));
        return RH_OK;
    }

    int W = w, H = h;
    bool down2 = false, resize = false;
    if (w > 512 || h > 512) {  // pdqhash.rs:181-188
        target_dimensions(w, h, 512, &W, &H);
        if (W * 2 == w && H * 2 == h)
            down2 = true;      // two rounded halving passes, fused into the front end
        else
            resize = true;     // general fixed-point Box convolution, run as its own kernels
    }
    DeviceBox bx{}, by{};
    if (resize) {
        BoxCoeffs cx, cy;
        box_precompute(w, W, &cx);
        box_precompute(h, H, &cy);
        RH_TRY(upload_box(ctx, cx, S_W14, &bx));
        RH_TRY(upload_box(ctx, cy, S_W15, &by));
    }
    const float *d_dct;
    RH_TRY(ensure_dct(ctx, &d_dct));
    TailOut out{o_hash.dev, o_q.dev, o_c.dev, o_dih.dev};
    const bool on_device = is_device_ptr(pixels);
    // host input is staged into 256-byte aligned buffers, so only the pitches matter there; a Box-resized
    // plane is written with its rows padded to 16 bytes (Wp), so it always qualifies
    const size_t Wp = ((size_t)W + 15) & ~(size_t)15;
    const bool fused = pdq_fused_supported(W, H) != 0 && !ctx->pdq_force_generic &&
                       (resize ? true
                               : pdq_fused_aligned(on_device ? (const void *)pixels : nullptr, row_pitch, img_pitch) != 0);
    // every other plane up to 512 x 512 whose width is a multiple of 8 (portrait photos, small images): the
    // float-chain fused kernel; what remains (odd widths, tiny planes, unaligned rows) takes the generic pipeline
    const bool fused_float = !fused && pdq_float_supported(W, H) != 0 && !ctx->pdq_force_generic &&
                             (resize ? true
                                     : pdq_fused_aligned(on_device ? (const void *)pixels : nullptr, row_pitch, img_pitch) != 0);
    // chunking: the generic pipeline keeps two f32 planes per image in scratch; host input is
    // streamed through two device buffers so the H2D copy of chunk k+1 overlaps the kernels of k.
    // (device-resident input to the fused kernel needs 9 KB of scratch per image: one launch for up
    // to 16384 images, so that the persistent CTAs see a long queue)
    // (the generic pipeline keeps ~10 B of scratch per plane pixel; 1024 images give its one-warp-per-32-columns
    // walks enough warps to hide their load latency)
    int64_t chunk = (fused || fused_float) ? (on_device ? 16384 : 2048) : 1024;
    if (resize) {   // full-resolution luma + the horizontally resized plane live in scratch
        int64_t c3 = (int64_t)((size_t(1) << 30) / ((size_t)w * h + (size_t)W * h + (size_t)W * H));
        if (c3 < 1) c3 = 1;
        if (chunk > c3) chunk = c3;
    }
    const size_t in_bytes_per = img_pitch;
    if (!on_device) {
        const size_t budget = size_t(512) << 20;
        int64_t c2 = (int64_t)(budget / in_bytes_per);
        if (c2 < 1) c2 = 1;
        if (chunk > c2) chunk = c2;
    }
    if (chunk > n) chunk = n;

    uint8_t *stage[2] = {nullptr, nullptr};
    if (!on_device) {
        void *p;
        RH_TRY(scratch(ctx, S_IN0, (size_t)chunk * in_bytes_per, &p));
        stage[0] = (uint8_t *)p;
        RH_TRY(scratch(ctx, S_IN1, (size_t)chunk * in_bytes_per, &p));
        stage[1] = (uint8_t *)p;
    }
    RH_CUDA(ctx, cudaEventRecord(ctx->ev_a, st));
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t cn = (n - off < chunk) ? (n - off) : chunk;
        const uint8_t *d_px = pixels + (size_t)off * img_pitch;
        int b = 0;
        if (!on_device) {
            // the two staging buffers alternate across chunks AND across calls (asynchronous calls overlap
            // the copy of the next batch with the kernels of this one); a buffer may be overwritten only once
            // the kernels of the chunk staged two chunks ago are done
            const uint64_t seq = ctx->stage_seq++;
            b = (int)(seq & 1);
            if (seq >= 2) RH_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
            size_t bytes = (size_t)(cn - 1) * img_pitch + row_pitch * (size_t)(h - 1) + (size_t)w * ch;
            RH_CUDA(ctx, cudaMemcpyAsync(stage[b], d_px, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            RH_CUDA(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
            RH_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_copy[b], 0));
            d_px = stage[b];
        }
        if (resize) {
            // luma at full resolution, Box-resize it (horizontal, then vertical), hash the planes
            void *p;
            RH_TRY(scratch(ctx, S_W5, (size_t)cn * w * h, &p));
            uint8_t *Lfull = (uint8_t *)p;
            RH_TRY(scratch(ctx, S_W6, (size_t)cn * W * h, &p));
            uint8_t *Lh = (uint8_t *)p;
            RH_TRY(scratch(ctx, S_W7, (size_t)cn * Wp * H, &p));
            uint8_t *Lr = (uint8_t *)p;
            if (layout == RH_LAYOUT_RGB8)
                RH_TRY(launch_luma<RH_LAYOUT_RGB8>(ctx, false, d_px, row_pitch, img_pitch, (int)cn, w, h, Lfull));
            else if (layout == RH_LAYOUT_RGBA8)
                RH_TRY(launch_luma<RH_LAYOUT_RGBA8>(ctx, false, d_px, row_pitch, img_pitch, (int)cn, w, h, Lfull));
            else
                RH_TRY(launch_luma<RH_LAYOUT_LUMA8>(ctx, false, d_px, row_pitch, img_pitch, (int)cn, w, h, Lfull));
            box_resize_kernel<<<cdiv((size_t)cn * W * h, 256), 256, 0, st>>>(Lfull, Lh, (size_t)cn, W, h, (size_t)w * h, w, 1,
                                                                            false, W, (size_t)W * h, bx.xmin, bx.cnt, bx.k, bx.ksize, bx.precision);
            RH_LAUNCHED(ctx, "box_resize_kernel");
            box_resize_kernel<<<cdiv((size_t)cn * W * H, 256), 256, 0, st>>>(Lh, Lr, (size_t)cn, W, H, (size_t)W * h, W, W,
                                                                            true, (int)Wp, Wp * H, by.xmin, by.cnt, by.k, by.ksize, by.precision);
            RH_LAUNCHED(ctx, "box_resize_kernel");
            if (fused)
                RH_TRY(pdq_fused_run(ctx, Lr, RH_LAYOUT_LUMA8, false, cn, W, H, Wp, Wp * H, out, off, d_dct));
            else if (fused_float)
                RH_TRY(pdq_float_run(ctx, Lr, RH_LAYOUT_LUMA8, false, cn, W, H, Wp, Wp * H, out, off, d_dct));
            else
                RH_TRY(generic_chunk(ctx, Lr, RH_LAYOUT_LUMA8, false, (int)cn, W, H, Wp, Wp * H, out, off, d_dct));
        } else if (fused)
            RH_TRY(pdq_fused_run(ctx, d_px, layout, down2, cn, W, H, row_pitch, img_pitch, out, off, d_dct));
        else if (fused_float)
            RH_TRY(pdq_float_run(ctx, d_px, layout, down2, cn, W, H, row_pitch, img_pitch, out, off, d_dct));
        else
            RH_TRY(generic_chunk(ctx, d_px, layout, down2, (int)cn, W, H, row_pitch, img_pitch, out, off, d_dct));
        if (!on_device) RH_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
    }
    RH_CUDA(ctx, cudaEventRecord(ctx->ev_b, st));
    if (o_valid.dev) {
        fill_u8_kernel<<<cdiv(n, 256), 256, 0, st>>>(o_valid.dev, (size_t)n, 1);
        RH_LAUNCHED(ctx, "fill_u8_kernel");
    }
    RH_TRY(o_hash.finish(ctx));
    RH_TRY(o_q.finish(ctx));
    RH_TRY(o_c.finish(ctx));
    RH_TRY(o_dih.finish(ctx));
    RH_TRY(o_valid.finish(ctx));
    ctx->last_units = (double)n;
    if (!wait) return RH_OK;
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    RH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    ctx->last_ms = ms;
    return RH_OK;
}

extern "C" {

int rh_pdq_hash_batch(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h, size_t row_pitch,
                      size_t img_pitch, uint8_t *out_hash, float *out_quality, float *out_coeffs,
                      uint8_t *out_dihedral, uint8_t *out_valid) {
    return pdq_hash_batch_impl(ctx, pixels, layout, n, w, h, row_pitch, img_pitch, out_hash, out_quality, out_coeffs,
                               out_dihedral, out_valid, true);
}

int rh_pdq_hash_batch_async(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h, size_t row_pitch,
                            size_t img_pitch, uint8_t *out_hash, float *out_quality, float *out_coeffs,
                            uint8_t *out_dihedral, uint8_t *out_valid, uint64_t *ticket) {
    int s = pdq_hash_batch_impl(ctx, pixels, layout, n, w, h, row_pitch, img_pitch, out_hash, out_quality, out_coeffs,
                                out_dihedral, out_valid, false);
    if (s != RH_OK) return s;
    // completion ticket: an event after everything this call queued (results included)
    const uint64_t t = ctx->tickets_issued++;
    cudaEvent_t &ev = ctx->ev_ticket[t % rh_ctx::kTickets];
    if (!ev) RH_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RH_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    if (ticket) *ticket = t;
    return RH_OK;
}

int rh_ctx_wait(rh_ctx *ctx, uint64_t ticket) {
    if (!ctx) return RH_EINVAL;
    if (ticket >= ctx->tickets_issued) return rh::fail(ctx, RH_EINVAL, "rh_ctx_wait: unknown ticket");
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ticket + rh_ctx::kTickets < ctx->tickets_issued) {
        // its event slot has been reused by a later call on the same stream: wait for that one instead
        RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RH_OK;
    }
    RH_CUDA(ctx, cudaEventSynchronize(ctx->ev_ticket[ticket % rh_ctx::kTickets]));
    return RH_OK;
}

int rh_pdq_from_buffer64(rh_ctx *ctx, const float *buf64x64, int64_t n, uint8_t *out_hash, float *out_quality,
                         float *out_coeffs, uint8_t *out_dihedral) {
    if (!ctx) return RH_EINVAL;
    if (n < 0 || (n > 0 && !buf64x64)) return fail(ctx, RH_EINVAL, "rh_pdq_from_buffer64: bad arguments");
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const float *d_in;
    RH_TRY(stage_in(ctx, buf64x64, (size_t)n * 4096, S_IN0, &d_in));
    OutBuf<uint8_t> o_hash, o_dih;
    OutBuf<float> o_q, o_c;
    RH_TRY(o_hash.prepare(ctx, out_hash, (size_t)n * 32, S_OUT0));
    RH_TRY(o_q.prepare(ctx, out_quality, (size_t)n, S_OUT1));
    RH_TRY(o_c.prepare(ctx, out_coeffs, (size_t)n * 256, S_OUT2));
    RH_TRY(o_dih.prepare(ctx, out_dihedral, (size_t)n * 256, S_OUT3));
    const float *d_dct;
    RH_TRY(ensure_dct(ctx, &d_dct));
    TailOut out{o_hash.dev, o_q.dev, o_c.dev, o_dih.dev};
    pdq_tail_kernel<SRC_BUF64><<<(unsigned)n, TAIL_THREADS, 0, st>>>(d_in, 64, 64, d_dct, out, 0);
    RH_LAUNCHED(ctx, "pdq_tail_kernel");
    RH_TRY(o_hash.finish(ctx));
    RH_TRY(o_q.finish(ctx));
    RH_TRY(o_c.finish(ctx));
    RH_TRY(o_dih.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}

static int from_coeffs(rh_ctx *ctx, const float *coeffs, int64_t n, uint8_t *out_hash, uint8_t *out_dihedral) {
    if (!ctx) return RH_EINVAL;
    if (n < 0 || (n > 0 && !coeffs)) return fail(ctx, RH_EINVAL, "from_coeffs: bad arguments");
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const float *d_in;
    RH_TRY(stage_in(ctx, coeffs, (size_t)n * 256, S_IN0, &d_in));
    OutBuf<uint8_t> o_hash, o_dih;
    RH_TRY(o_hash.prepare(ctx, out_hash, (size_t)n * 32, S_OUT0));
    RH_TRY(o_dih.prepare(ctx, out_dihedral, (size_t)n * 256, S_OUT3));
    TailOut out{o_hash.dev, nullptr, nullptr, o_dih.dev};
    pdq_tail_kernel<SRC_COEFFS><<<(unsigned)n, TAIL_THREADS, 0, st>>>(d_in, 0, 0, nullptr, out, 0);
    RH_LAUNCHED(ctx, "pdq_tail_kernel");
    RH_TRY(o_hash.finish(ctx));
    RH_TRY(o_dih.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}

int rh_pdq_hash_from_coeffs(rh_ctx *ctx, const float *coeffs, int64_t n, uint8_t *out_hash) {
    return from_coeffs(ctx, coeffs, n, out_hash, nullptr);
}

int rh_pdq_dihedral_from_coeffs(rh_ctx *ctx, const float *coeffs, int64_t n, uint8_t *out_dihedral) {
    return from_coeffs(ctx, coeffs, n, nullptr, out_dihedral);
}

}  // extern "C"
