// phash.cu -- 64-bit DCT pHash, DctPhash::hash_image (phash.rs:48-83) over a batch.
//
//   img.resize_exact(32, 32, Triangle)   in the image's native colour type   (phash.rs:51-52)
//   .to_luma8()                           Rec.709 integer luma                (phash.rs:53)
//   2-D DCT-II of the 32 x 32 plane, top-left 8 x 8                          (phash.rs:95-128)
//   median of the 63 non-DC values, bit 63-i = low[i] > median               (phash.rs:67-80)
//
// The resize / luma / DCT arithmetic lives in the `image` and `rustdct` crates, which are not in
// the reference tree: this kernel follows the oracle's restatement of them (oracle_phash.c) step
// for step -- vertical pass into f32, horizontal pass with clamp + round-half-away, naive DCT
// order with an f64-computed cosine table, separate mul and add.  PARITY UNPINNED with respect to
// the real crates; exact with respect to the oracle.
// One CTA of 256 threads per image.  Not a throughput target: phdupes itself never calls pHash.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

using namespace rh;

constexpr int PH = 32;          // phash.rs:20 DCT_SIZE
constexpr int PH_THREADS = 256;
constexpr int PH_MAXW = 4096;   // widest image row the kernel stages (floats: w * ch <= 16384)

struct AxisWeights {            // per output coordinate: first tap, tap count, offset into ws[]
    int left[PH], cnt[PH], off[PH];
};

struct PhashArgs {
    const uint8_t *px;
    size_t row_pitch, img_pitch;
    int w, h;
    AxisWeights v, hz;          // vertical (over rows) and horizontal (over columns) taps
    const float *ws_v, *ws_h;   // normalised Triangle weights
    const float *cs;            // cos(pi (n + 1/2) k / 32), [k][n]
    uint64_t *out_hash, *out_dihedral;
};

// phash.rs:150-255 on the device (same bit rules as ctx.cu's host versions)
__device__ __forceinline__ uint64_t dev_bit(uint64_t h, int x, int y) { return (h >> (63 - (8 * y + x))) & 1ull; }
__device__ uint64_t dev_transform(uint64_t h, int kind) {   // 0 rot90, 1 rot180, 2 rot270, 3 flip
    uint64_t r = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            uint64_t b;
            if (kind == 0) b = dev_bit(h, y, x) ^ (uint64_t)(x & 1);
            else if (kind == 1) b = dev_bit(h, x, y) ^ (uint64_t)((x + y) & 1);
            else if (kind == 2) b = dev_bit(h, y, x) ^ (uint64_t)(y & 1);
            else b = dev_bit(h, x, y) ^ (uint64_t)(x & 1);
            r |= b << (63 - (8 * y + x));
        }
    return r;
}

template <int CH>
__global__ void __launch_bounds__(PH_THREADS) phash_kernel(const PhashArgs a) {
    extern __shared__ float s_row[];            // one vertically sampled row: w * CH floats
    __shared__ uint8_t s_small[PH * PH * 4];    // the 32 x 32 resized image
    __shared__ uint8_t s_luma[PH * PH];
    __shared__ float s_rows[PH][PH + 1];        // row DCTs
    __shared__ float s_low[64];
    __shared__ float s_median;
    const uint8_t *img = a.px + (size_t)blockIdx.x * a.img_pitch;
    const int w = a.w, h = a.h, tid = threadIdx.x;

    if (w == PH && h == PH) {
        for (int i = tid; i < PH * PH * CH; i += PH_THREADS) s_small[i] = img[(size_t)(i / (PH * CH)) * a.row_pitch + i % (PH * CH)];
    } else {
        for (int oy = 0; oy < PH; oy++) {
            // vertical_sample: f32, taps ascending, mul then add
            const int left = a.v.left[oy], cnt = a.v.cnt[oy];
            const float *ws = a.ws_v + a.v.off[oy];
            for (int xc = tid; xc < w * CH; xc += PH_THREADS) {
                float t = 0.0f;
                for (int i = 0; i < cnt; i++)
                    t = __fadd_rn(t, __fmul_rn((float)img[(size_t)(left + i) * a.row_pitch + xc], ws[i]));
                s_row[xc] = t;
            }
            __syncthreads();
            // horizontal_sample for this row: clamp to [0, 255], round half away from zero
            for (int o = tid; o < PH * CH; o += PH_THREADS) {
                const int ox = o / CH, c = o % CH;
                const int l2 = a.hz.left[ox], n2 = a.hz.cnt[ox];
                const float *w2 = a.ws_h + a.hz.off[ox];
                float t = 0.0f;
                for (int i = 0; i < n2; i++) t = __fadd_rn(t, __fmul_rn(s_row[(l2 + i) * CH + c], w2[i]));
                t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
                s_small[(oy * PH + ox) * CH + c] = (uint8_t)roundf(t);
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // to_luma8: Rec.709 integer, truncating (image 0.25)
    for (int i = tid; i < PH * PH; i += PH_THREADS) {
        if (CH == 1) s_luma[i] = s_small[i];
        else {
            const uint8_t *p = s_small + i * CH;
            s_luma[i] = (uint8_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u);
        }
    }
    __syncthreads();
    // row DCTs: rows[y][k] = sum_n luma[y][n] * cs[k][n], n ascending
    for (int o = tid; o < PH * PH; o += PH_THREADS) {
        const int y = o / PH, k = o % PH;
        float s = 0.0f;
        for (int n = 0; n < PH; n++) s = __fadd_rn(s, __fmul_rn((float)s_luma[y * PH + n], a.cs[k * PH + n]));
        s_rows[y][k] = s;
    }
    __syncthreads();
    // column DCTs of the 8 x 8 corner (crop_8x8, phash.rs:121-128)
    if (tid < 64) {
        const int ky = tid >> 3, kx = tid & 7;
        float s = 0.0f;
        for (int n = 0; n < PH; n++) s = __fadd_rn(s, __fmul_rn(s_rows[n][kx], a.cs[ky * PH + n]));
        s_low[tid] = s;
    }
    __syncthreads();
    // median = sorted[31] of the 63 non-DC values (phash.rs:67-71)
    if (tid >= 1 && tid < 64) {
        const float v = s_low[tid];
        int rank = 0;
        for (int j = 1; j < 64; j++) {
            const float u = s_low[j];
            rank += (u < v) || (u == v && j < tid);
        }
        if (rank == 31) s_median = v;
    }
    __syncthreads();
    if (tid < 32) {
        // bit 63 - i = low[i] > median (phash.rs:74-80); lanes hold i and i + 32
        const uint32_t hi = __ballot_sync(0xFFFFFFFFu, s_low[tid] > s_median);        // i = 0..31
        const uint32_t lo = __ballot_sync(0xFFFFFFFFu, s_low[tid + 32] > s_median);   // i = 32..63
        if (tid == 0) {
            const uint64_t hash = ((uint64_t)__brev(hi) << 32) | (uint64_t)__brev(lo);
            if (a.out_hash) a.out_hash[blockIdx.x] = hash;
            if (a.out_dihedral) {   // phash.rs:242-255
                uint64_t *d = a.out_dihedral + (size_t)blockIdx.x * 8;
                const uint64_t f = dev_transform(hash, 3);
                d[0] = hash;
                d[1] = dev_transform(hash, 0);
                d[2] = dev_transform(hash, 1);
                d[3] = dev_transform(hash, 2);
                d[4] = f;
                d[5] = dev_transform(f, 0);
                d[6] = dev_transform(f, 1);
                d[7] = dev_transform(f, 2);
            }
        }
    }
}

// image 0.25 imageops::resize(.., Triangle) tap table for one axis (oracle_phash.c
// sample_axis_weights, same f32 steps)
void axis_weights(int in, AxisWeights *aw, std::vector<float> *ws) {
    ws->clear();
    const float ratio = (float)in / (float)PH;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 1.0f * sratio;
    for (int o = 0; o < PH; o++) {
        volatile float c = ((float)o + 0.5f) * ratio;
        long left = (long)floorf(c - support);
        if (left < 0) left = 0;
        if (left > (long)in - 1) left = (long)in - 1;
        long right = (long)ceilf(c + support);
        if (right < left + 1) right = left + 1;
        if (right > (long)in) right = (long)in;
        c = c - 0.5f;
        const int cnt = (int)(right - left);
        aw->left[o] = (int)left;
        aw->cnt[o] = cnt;
        aw->off[o] = (int)ws->size();
        volatile float sum = 0.0f;
        for (int i = 0; i < cnt; i++) {
            volatile float x = ((float)(left + i) - c) / sratio;
            const float ax = fabsf(x);
            const float wgt = ax < 1.0f ? 1.0f - ax : 0.0f;
            ws->push_back(wgt);
            sum += wgt;
        }
        for (int i = 0; i < cnt; i++) (*ws)[aw->off[o] + i] /= sum;
    }
}

}  // namespace

extern "C" int rh_phash_batch(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h,
                              size_t row_pitch, size_t img_pitch, uint64_t *out_hash, uint64_t *out_dihedral) {
    if (!ctx) return RH_EINVAL;
    if (n < 0 || w <= 0 || h <= 0 || (n > 0 && !pixels)) return fail(ctx, RH_EINVAL, "rh_phash_batch: bad arguments");
    if (layout != RH_LAYOUT_RGB8 && layout != RH_LAYOUT_RGBA8 && layout != RH_LAYOUT_LUMA8)
        return fail(ctx, RH_EINVAL, "rh_phash_batch: unknown layout");
    const int ch = layout == RH_LAYOUT_RGB8 ? 3 : (layout == RH_LAYOUT_RGBA8 ? 4 : 1);
    if ((size_t)w * ch > (size_t)PH_MAXW * 4) return fail(ctx, RH_EUNSUPPORTED, "rh_phash_batch: image wider than 16384 samples per row");
    if (row_pitch == 0) row_pitch = (size_t)w * ch;
    if (img_pitch == 0) img_pitch = row_pitch * h;
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    const uint8_t *d_px;
    RH_TRY(stage_in(ctx, pixels, (size_t)n * img_pitch, S_IN0, &d_px));
    OutBuf<uint64_t> o_hash, o_dih;
    RH_TRY(o_hash.prepare(ctx, out_hash, (size_t)n, S_OUT0));
    RH_TRY(o_dih.prepare(ctx, out_dihedral, (size_t)n * 8, S_OUT1));

    PhashArgs a;
    std::vector<float> wv, wh, cs((size_t)PH * PH);
    axis_weights(h, &a.v, &wv);
    axis_weights(w, &a.hz, &wh);
    for (int k = 0; k < PH; k++)
        for (int m = 0; m < PH; m++) cs[(size_t)k * PH + m] = (float)cos(M_PI * (m + 0.5) * k / PH);
    void *p;
    RH_TRY(scratch(ctx, S_W0, (wv.size() + wh.size() + cs.size()) * sizeof(float), &p));
    float *d_tab = (float *)p;
    std::vector<float> tab;
    tab.insert(tab.end(), wv.begin(), wv.end());
    tab.insert(tab.end(), wh.begin(), wh.end());
    tab.insert(tab.end(), cs.begin(), cs.end());
    RH_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    RH_CUDA(ctx, cudaStreamSynchronize(st));   // `tab` is a stack object
    a.px = d_px;
    a.row_pitch = row_pitch;
    a.img_pitch = img_pitch;
    a.w = w;
    a.h = h;
    a.ws_v = d_tab;
    a.ws_h = d_tab + wv.size();
    a.cs = d_tab + wv.size() + wh.size();
    a.out_hash = o_hash.dev;
    a.out_dihedral = o_dih.dev;
    const size_t smem = (size_t)w * ch * sizeof(float);
    if (ch == 3) {
        RH_CUDA(ctx, cudaFuncSetAttribute(phash_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        phash_kernel<3><<<(unsigned)n, PH_THREADS, smem, st>>>(a);
    } else if (ch == 4) {
        RH_CUDA(ctx, cudaFuncSetAttribute(phash_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        phash_kernel<4><<<(unsigned)n, PH_THREADS, smem, st>>>(a);
    } else {
        RH_CUDA(ctx, cudaFuncSetAttribute(phash_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        phash_kernel<1><<<(unsigned)n, PH_THREADS, smem, st>>>(a);
    }
    RH_LAUNCHED(ctx, "phash_kernel");
    RH_TRY(o_hash.finish(ctx));
    RH_TRY(o_dih.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}
