/*
 * oracle/oracle_phash.c -- CPU restatement of /root/reference/src/phash.rs.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The bit-level functions (phash.rs:137-255) are exactly portable.  The image
 * path (phash.rs:48-83) goes through the `image` 0.25.10 and `rustdct` 0.7.1 crates,
 * neither of which is under /root/reference: resize/luma are restated from memory
 * and the DCT uses a fixed naive f32 order instead of rustdct's split-radix
 * butterflies => PARITY UNPINNED for orc_phash_image / orc_phash_from_luma32
 * (bits whose coefficient sits within a few ulp of the median may differ).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DCT_SIZE 32 /* phash.rs:20 */
#define HASH_SIZE 8 /* phash.rs:21 */

/* phash.rs:150-171 */
uint64_t orc_phash_rot90(uint64_t hash) {
    uint64_t result = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int src_idx = 8 * y + x;
            int dst_x = y, dst_y = x;
            int dst_idx = 8 * dst_y + dst_x;
            uint64_t bit = (hash >> (63 - src_idx)) & 1;
            if (dst_x % 2 != 0) bit ^= 1;
            result |= bit << (63 - dst_idx);
        }
    return result;
}

/* phash.rs:175-188 */
uint64_t orc_phash_rot180(uint64_t hash) {
    uint64_t result = 0;
    for (int i = 0; i < 64; i++) {
        int x = i % 8, y = i / 8;
        uint64_t bit = (hash >> (63 - i)) & 1;
        if ((x + y) % 2 != 0) bit ^= 1;
        result |= bit << (63 - i);
    }
    return result;
}

/* phash.rs:191-212 */
uint64_t orc_phash_rot270(uint64_t hash) {
    uint64_t result = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int src_idx = 8 * y + x;
            int dst_x = y, dst_y = x;
            int dst_idx = 8 * dst_y + dst_x;
            uint64_t bit = (hash >> (63 - src_idx)) & 1;
            if (dst_y % 2 != 0) bit ^= 1;
            result |= bit << (63 - dst_idx);
        }
    return result;
}

/* phash.rs:220-230 */
uint64_t orc_phash_flip_h(uint64_t hash) {
    uint64_t result = 0;
    for (int i = 0; i < 64; i++) {
        int x = i % 8;
        uint64_t bit = (hash >> (63 - i)) & 1;
        if (x % 2 != 0) bit ^= 1;
        result |= bit << (63 - i);
    }
    return result;
}

/* phash.rs:242-255 */
void orc_phash_dihedral(uint64_t h, uint64_t *out) {
    uint64_t f = orc_phash_flip_h(h);
    out[0] = h;
    out[1] = orc_phash_rot90(h);
    out[2] = orc_phash_rot180(h);
    out[3] = orc_phash_rot270(h);
    out[4] = f;
    out[5] = orc_phash_rot90(f);
    out[6] = orc_phash_rot180(f);
    out[7] = orc_phash_rot270(f);
}

/* phash.rs:137-143 */
uint64_t orc_phash_rot_invariant(uint64_t h) {
    uint64_t a = orc_phash_rot90(h), b = orc_phash_rot180(h), c = orc_phash_rot270(h);
    uint64_t m = h;
    if (a < m) m = a;
    if (b < m) m = b;
    if (c < m) m = c;
    return m;
}

static int cmp_f32(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return x < y ? -1 : (x > y ? 1 : 0); /* partial_cmp, no NaNs possible here */
}

/*
 * phash.rs:55-83 + 95-128.  rustdct's DCT-II is unnormalised:
 *   X_k = sum_n x_n cos(pi (n + 1/2) k / 32).
 * Order used here (and mirrored by the CUDA kernel): cos table computed in f64 and
 * rounded to f32; rows first, n ascending, separate mul/add in f32; then columns.
 */
uint64_t orc_phash_from_luma32(const uint8_t *luma) {
    static float cs[DCT_SIZE][DCT_SIZE];
    static int init = 0;
    if (!init) {
        for (int k = 0; k < DCT_SIZE; k++)
            for (int n = 0; n < DCT_SIZE; n++) cs[k][n] = (float)cos(M_PI * (n + 0.5) * k / DCT_SIZE);
        init = 1;
    }
    float rows[DCT_SIZE][DCT_SIZE];
    for (int y = 0; y < DCT_SIZE; y++)
        for (int k = 0; k < DCT_SIZE; k++) {
            float s = 0.0f;
            for (int n = 0; n < DCT_SIZE; n++) {
                float p = (float)luma[y * DCT_SIZE + n] * cs[k][n];
                s = s + p;
            }
            rows[y][k] = s;
        }
    float low[HASH_SIZE * HASH_SIZE]; /* crop_8x8 :121-128 */
    for (int ky = 0; ky < HASH_SIZE; ky++)
        for (int kx = 0; kx < HASH_SIZE; kx++) {
            float s = 0.0f;
            for (int n = 0; n < DCT_SIZE; n++) {
                float p = rows[n][kx] * cs[ky][n];
                s = s + p;
            }
            low[ky * HASH_SIZE + kx] = s;
        }
    float sorted[63]; /* :67-71: DC removed, sorted[63/2 = 31] */
    memcpy(sorted, low + 1, sizeof(sorted));
    qsort(sorted, 63, sizeof(float), cmp_f32);
    float median = sorted[31];
    uint64_t hash = 0; /* :74-80 */
    for (int i = 0; i < 64; i++)
        if (low[i] > median) hash |= 1ull << (63 - i);
    return hash;
}

/* image 0.25 imageops::resize(.., Triangle) restated from memory (UNVERIFIED):
 * vertical pass into f32 (no rounding), then horizontal pass, clamp, round half
 * away from zero. */
static float tri(float x) {
    float a = fabsf(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}

static void sample_axis_weights(uint32_t in, uint32_t out, uint32_t o, uint32_t *left_out, uint32_t *cnt_out, float *ws) {
    float ratio = (float)in / (float)out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = 1.0f * sratio;
    float c = ((float)o + 0.5f) * ratio;
    long left = (long)floorf(c - support);
    if (left < 0) left = 0;
    if (left > (long)in - 1) left = (long)in - 1;
    long right = (long)ceilf(c + support);
    if (right < left + 1) right = left + 1;
    if (right > (long)in) right = (long)in;
    c = c - 0.5f;
    float sum = 0.0f;
    uint32_t cnt = (uint32_t)(right - left);
    for (uint32_t i = 0; i < cnt; i++) {
        float w = tri(((float)(left + i) - c) / sratio);
        ws[i] = w;
        sum += w;
    }
    for (uint32_t i = 0; i < cnt; i++) ws[i] /= sum;
    *left_out = (uint32_t)left;
    *cnt_out = cnt;
}

uint64_t orc_phash_image(const uint8_t *px, int layout, uint32_t w, uint32_t h, uint8_t *luma32_out) {
    int ch = layout == ORC_LAYOUT_LUMA8 ? 1 : (layout == ORC_LAYOUT_RGBA8 ? 4 : 3);
    uint8_t small[DCT_SIZE * DCT_SIZE * 4];
    if (w == DCT_SIZE && h == DCT_SIZE) {
        memcpy(small, px, (size_t)DCT_SIZE * DCT_SIZE * ch);
    } else {
        float *ws = (float *)malloc(sizeof(float) * ((size_t)(w > h ? w : h) + 4));
        float *tmp = (float *)malloc(sizeof(float) * (size_t)w * DCT_SIZE * ch);
        for (uint32_t oy = 0; oy < DCT_SIZE; oy++) { /* vertical_sample */
            uint32_t left, cnt;
            sample_axis_weights(h, DCT_SIZE, oy, &left, &cnt, ws);
            for (uint32_t x = 0; x < w; x++)
                for (int c = 0; c < ch; c++) {
                    float t = 0.0f;
                    for (uint32_t i = 0; i < cnt; i++) {
                        float p = (float)px[((size_t)(left + i) * w + x) * ch + c] * ws[i];
                        t = t + p;
                    }
                    tmp[((size_t)oy * w + x) * ch + c] = t;
                }
        }
        for (uint32_t ox = 0; ox < DCT_SIZE; ox++) { /* horizontal_sample */
            uint32_t left, cnt;
            sample_axis_weights(w, DCT_SIZE, ox, &left, &cnt, ws);
            for (uint32_t y = 0; y < DCT_SIZE; y++)
                for (int c = 0; c < ch; c++) {
                    float t = 0.0f;
                    for (uint32_t i = 0; i < cnt; i++) {
                        float p = tmp[((size_t)y * w + left + i) * ch + c] * ws[i];
                        t = t + p;
                    }
                    if (t < 0.0f) t = 0.0f;
                    if (t > 255.0f) t = 255.0f;
                    small[((size_t)y * DCT_SIZE + ox) * ch + c] = (uint8_t)roundf(t);
                }
        }
        free(ws);
        free(tmp);
    }
    uint8_t luma[DCT_SIZE * DCT_SIZE];
    for (int i = 0; i < DCT_SIZE * DCT_SIZE; i++) {
        if (ch == 1) luma[i] = small[i];
        else { /* image crate to_luma8: Rec.709 integer, truncating */
            const uint8_t *p = small + (size_t)i * ch;
            luma[i] = (uint8_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u);
        }
    }
    if (luma32_out) memcpy(luma32_out, luma, sizeof(luma));
    return orc_phash_from_luma32(luma);
}
