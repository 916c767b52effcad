/*
 * oracle/oracle_pdq.c -- CPU restatement of /root/reference/src/pdqhash.rs.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Compile with -ffp-contract=off.
 *
 * Every function cites the reference lines it follows.  The arithmetic is f32
 * with separate multiply and add (Rust never contracts to FMA) and the exact
 * accumulation orders of the reference.
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Per-thread scratch planes, grown on demand and reused across images -- the analogue of the
 * reference's thread-local Resizer (pdqhash.rs:38-42).  Without it every image mmaps and
 * unmaps ~3 MB and the page-fault / TLB-shootdown traffic serialises the worker threads. */
enum { SCR_LUMA = 0, SCR_RESIZED, SCR_RTMP, SCR_PLANE, SCR_JTMP, SCR_COUNT };
static __thread void *t_scr[SCR_COUNT];
static __thread size_t t_scr_bytes[SCR_COUNT];
static void *scratch(int slot, size_t bytes) {
    if (t_scr_bytes[slot] < bytes) {
        free(t_scr[slot]);
        t_scr[slot] = malloc(bytes);
        t_scr_bytes[slot] = t_scr[slot] ? bytes : 0;
    }
    return t_scr[slot];
}
static void scratch_release(void) {
    for (int i = 0; i < SCR_COUNT; i++) {
        free(t_scr[i]);
        t_scr[i] = NULL;
        t_scr_bytes[i] = 0;
    }
}

#define MIN_HASHABLE_DIM 5u   /* pdqhash.rs:17 */
#define JAROSZ_PASSES 2       /* pdqhash.rs:18 */
#define DOWNSAMPLE_DIMS 512u  /* pdqhash.rs:19 */
#define BUF 64                /* pdqhash.rs:20 */
#define DCT_OUT 16            /* pdqhash.rs:21 */
#define DCT_FREQ_OFFSET 1     /* pdqhash.rs:31 */

/* pdqhash.rs:224-235 */
void orc_target_dimensions(uint32_t w, uint32_t h, uint32_t max_dim, uint32_t *nw, uint32_t *nh) {
    if (w == 0 || h == 0) {
        *nw = w > 1 ? w : 1;
        *nh = h > 1 ? h : 1;
        return;
    }
    if (w > h) {
        uint64_t v = (uint64_t)h * (uint64_t)max_dim / (uint64_t)w;
        *nw = max_dim;
        *nh = (uint32_t)(v > 1 ? v : 1);
    } else {
        uint64_t v = (uint64_t)w * (uint64_t)max_dim / (uint64_t)h;
        *nw = (uint32_t)(v > 1 ? v : 1);
        *nh = max_dim;
    }
}

/* pdqhash.rs:268-284: (299r + 587g + 114b + 500) / 1000, truncating u32 division.
 * RGBA ignores alpha; Luma8 is copied. */
void orc_luma601(const uint8_t *px, int layout, size_t n, uint8_t *out) {
    if (layout == ORC_LAYOUT_LUMA8) {
        memcpy(out, px, n);
        return;
    }
    size_t step = layout == ORC_LAYOUT_RGBA8 ? 4 : 3;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *p = px + i * step;
        out[i] = (uint8_t)((299u * p[0] + 587u * p[1] + 114u * p[2] + 500u) / 1000u);
    }
}

/*
 * pdqhash.rs:203-220 calls fast_image_resize 6.1.0 (Cargo.lock) with
 * ResizeAlg::Convolution(FilterType::Box) on a U8 plane.  That crate is not under
 * /root/reference; this is its published (Pillow-derived) algorithm restated from
 * memory -- UNVERIFIED against the crate:
 *   per axis: scale = in/out, support = 0.5*max(scale,1); for each output o:
 *   center = (o+0.5)*scale, xmin = max(0,(int)(center-support+0.5)),
 *   xmax = min(in,(int)(center+support+0.5)); box weight 1 inside (-0.5,0.5],
 *   normalised by their sum (f64); coefficients quantised to i16 at the largest
 *   precision p for which the largest coefficient stays < 2^15; each pass computes
 *   (sum(px*k) + (1 << (p-1))) >> p clipped to [0,255].  Horizontal pass first into
 *   a u8 temporary, vertical second.
 * For an exact 2x ratio this is ((a+b+1)>>1) per row pair, then the same
 * vertically, whatever the precision -- which is all BASELINE's configs need.
 */
typedef struct {
    int *xmin;
    int *cnt;
    int16_t *k;
    int ksize;
    int precision;
} box_coeffs;

static double box_filter(double x) { return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0; }

static int box_precompute(int in_size, int out_size, box_coeffs *bc) {
    double scale = (double)in_size / (double)out_size;
    double fscale = scale < 1.0 ? 1.0 : scale;
    double support = 0.5 * fscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double *pre = (double *)calloc((size_t)out_size * ksize, sizeof(double));
    bc->xmin = (int *)malloc(sizeof(int) * out_size);
    bc->cnt = (int *)malloc(sizeof(int) * out_size);
    bc->k = (int16_t *)calloc((size_t)out_size * ksize, sizeof(int16_t));
    bc->ksize = ksize;
    if (!pre || !bc->xmin || !bc->cnt || !bc->k) return -1;
    double ss = 1.0 / fscale;
    double maxk = 0.0;
    for (int o = 0; o < out_size; o++) {
        double center = (o + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        int cnt = xmax - xmin;
        double ww = 0.0;
        double *k = pre + (size_t)o * ksize;
        for (int x = 0; x < cnt; x++) {
            double w = box_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < cnt; x++) {
            if (ww != 0.0) k[x] /= ww;
            if (k[x] > maxk) maxk = k[x];
        }
        bc->xmin[o] = xmin;
        bc->cnt[o] = cnt;
    }
    int p;
    for (p = 0; p < 32 - 8 - 2; p++) {
        int next = (int)(0.5 + maxk * (double)(1 << (p + 1)));
        if (next >= (1 << 15)) break;
    }
    bc->precision = p;
    for (size_t i = 0; i < (size_t)out_size * ksize; i++) {
        double v = pre[i] * (double)(1 << p);
        bc->k[i] = (int16_t)(v < 0 ? (int)(v - 0.5) : (int)(v + 0.5));
    }
    free(pre);
    return 0;
}

static void box_free(box_coeffs *bc) {
    free(bc->xmin);
    free(bc->cnt);
    free(bc->k);
}

static inline uint8_t clip8(int v, int p) {
    v >>= p;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

int orc_resize_box_u8(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
    if (dw <= 0 || dh <= 0) return -1; /* pdqhash.rs:204-206 */
    box_coeffs bx, by;
    if (box_precompute(sw, dw, &bx)) return -1;
    if (box_precompute(sh, dh, &by)) return -1;
    uint8_t *tmp = (uint8_t *)scratch(SCR_RTMP, (size_t)dw * sh);
    if (!tmp) return -1;
    for (int y = 0; y < sh; y++) {
        const uint8_t *row = src + (size_t)y * sw;
        for (int o = 0; o < dw; o++) {
            const int16_t *k = bx.k + (size_t)o * bx.ksize;
            int acc = 1 << (bx.precision - 1);
            for (int x = 0; x < bx.cnt[o]; x++) acc += row[bx.xmin[o] + x] * k[x];
            tmp[(size_t)y * dw + o] = clip8(acc, bx.precision);
        }
    }
    for (int o = 0; o < dh; o++) {
        const int16_t *k = by.k + (size_t)o * by.ksize;
        for (int x = 0; x < dw; x++) {
            int acc = 1 << (by.precision - 1);
            for (int y = 0; y < by.cnt[o]; y++) acc += tmp[(size_t)(by.xmin[o] + y) * dw + x] * k[y];
            dst[(size_t)o * dw + x] = clip8(acc, by.precision);
        }
    }
    box_free(&bx);
    box_free(&by);
    return 0;
}

/* pdqhash.rs:341-396 */
void orc_box_one_d(const float *in, size_t in_start, float *out, size_t out_start, size_t len,
                   size_t stride, size_t win) {
    size_t lim = len > 1 ? len : 1;
    if (win < 1) win = 1;
    if (win > lim) win = lim; /* :351 */
    size_t half = (win + 2) / 2; /* :352 */
    size_t phase_1 = half - 1;
    size_t phase_2 = win - half + 1;
    size_t phase_3 = len > win ? len - win : 0;
    size_t phase_4 = half - 1;

    size_t li = in_start, ri = in_start, oi = out_start;
    float sum = 0.0f;
    float curr_win = 0.0f;

    for (size_t i = 0; i < phase_1; i++) { /* :366-370 */
        sum = sum + in[ri];
        curr_win += 1.0f;
        ri += stride;
    }
    for (size_t i = 0; i < phase_2; i++) { /* :372-378 */
        sum = sum + in[ri];
        curr_win += 1.0f;
        out[oi] = sum / curr_win;
        ri += stride;
        oi += stride;
    }
    for (size_t i = 0; i < phase_3; i++) { /* :380-387 */
        sum = sum + in[ri];
        sum = sum - in[li];
        out[oi] = sum / curr_win;
        li += stride;
        ri += stride;
        oi += stride;
    }
    for (size_t i = 0; i < phase_4; i++) { /* :389-395 */
        sum = sum - in[li];
        curr_win -= 1.0f;
        out[oi] = sum / curr_win;
        li += stride;
        oi += stride;
    }
}

/* pdqhash.rs:398-426 */
void orc_jarosz(float *buf, size_t rows, size_t cols, size_t w_rows, size_t w_cols, size_t nreps) {
    float *tmp = (float *)scratch(SCR_JTMP, rows * cols * sizeof(float));
    for (size_t rep = 0; rep < nreps; rep++) {
        for (size_t i = 0; i < rows; i++) orc_box_one_d(buf, i * cols, tmp, i * cols, cols, 1, w_rows);
        for (size_t j = 0; j < cols; j++) orc_box_one_d(tmp, j, buf, j, rows, cols, w_cols);
    }
}

/* pdqhash.rs:428-443 */
void orc_decimate64(const float *in, size_t in_r, size_t in_c, float *out) {
    for (size_t i = 0; i < BUF; i++) {
        size_t ini = ((i * 2 + 1) * in_r) / (BUF * 2);
        for (size_t j = 0; j < BUF; j++) {
            size_t inj = ((j * 2 + 1) * in_c) / (BUF * 2);
            out[i * BUF + j] = in[ini * in_c + inj];
        }
    }
}

/* pdqhash.rs:445-460: vertical pairs first, then horizontal, f32 running sum. */
float orc_quality(const float *buf, size_t R, size_t C) {
    float sum = 0.0f;
    for (size_t i = 0; i + 1 < R; i++)
        for (size_t j = 0; j < C; j++) {
            float a = buf[i * C + j], b = buf[(i + 1) * C + j];
            float d = a - b;
            d = d * 100.0f;
            d = d / 255.0f;
            sum = sum + truncf(fabsf(d));
        }
    for (size_t i = 0; i < R; i++)
        for (size_t j = 0; j + 1 < C; j++) {
            float a = buf[i * C + j], b = buf[i * C + j + 1];
            float d = a - b;
            d = d * 100.0f;
            d = d / 255.0f;
            sum = sum + truncf(fabsf(d));
        }
    float q = sum / 90.0f;
    return q > 1.0f ? 1.0f : q;
}

/* pdqhash.rs:287-304.  All f32; cos is the platform cosf (Rust f32::cos -> libm). */
void orc_dct_matrix(float *D) {
    const float PI_F = 3.14159265358979323846f; /* std::f32::consts::PI */
    float num_cols = (float)BUF;
    float inv_sqrt_cols = 1.0f / sqrtf(num_cols);
    float sqrt_2 = sqrtf(2.0f);
    for (int i = 0; i < DCT_OUT; i++) {
        float freq = (float)(i + DCT_FREQ_OFFSET);
        float norm = freq == 0.0f ? inv_sqrt_cols : inv_sqrt_cols * sqrt_2;
        for (int j = 0; j < BUF; j++) {
            float a = PI_F * freq;
            float b = 2.0f * (float)j + 1.0f;
            float num = a * b;
            float angle = num / (2.0f * num_cols);
            D[i * BUF + j] = norm * cosf(angle);
        }
    }
}

static float g_dct[DCT_OUT * BUF];
static pthread_once_t g_dct_once = PTHREAD_ONCE_INIT;
static void dct_init(void) { orc_dct_matrix(g_dct); }

/* pdqhash.rs:306-336 */
void orc_dct64_to_16(const float *in, float *out) {
    pthread_once(&g_dct_once, dct_init);
    float inter[DCT_OUT][BUF];
    memset(inter, 0, sizeof(inter));
    for (int i = 0; i < DCT_OUT; i++)
        for (int k = 0; k < BUF; k++) {
            float coeff = g_dct[i * BUF + k];
            for (int j = 0; j < BUF; j++) {
                float p = coeff * in[k * BUF + j];
                inter[i][j] = inter[i][j] + p;
            }
        }
    for (int i = 0; i < DCT_OUT; i++)
        for (int j = 0; j < DCT_OUT; j++) {
            float sum = 0.0f;
            for (int k = 0; k < BUF; k++) {
                float p = inter[i][k] * g_dct[j * BUF + k];
                sum = sum + p;
            }
            out[i * DCT_OUT + j] = sum;
        }
}

/* pdqhash.rs:127-137 */
static inline float apply_sign(float v, int r, int c, int neg_rows, int neg_cols) {
    int flip_r = neg_rows && ((r + DCT_FREQ_OFFSET) % 2 == 1);
    int flip_c = neg_cols && ((c + DCT_FREQ_OFFSET) % 2 == 1);
    return (flip_r ^ flip_c) ? -v : v;
}

/* f32::total_cmp as an integer key */
static inline int32_t total_key(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}
static int cmp_total(const void *a, const void *b) {
    int32_t x = total_key(*(const float *)a), y = total_key(*(const float *)b);
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* pdqhash.rs:116-124: element of rank (256-1)/2 = 127 under total_cmp */
static float coefficient_median(const float *c, int neg_rows, int neg_cols) {
    float buf[256];
    for (int idx = 0; idx < 256; idx++)
        buf[idx] = apply_sign(c[idx], idx / DCT_OUT, idx % DCT_OUT, neg_rows, neg_cols);
    qsort(buf, 256, sizeof(float), cmp_total);
    return buf[127];
}

/* pdqhash.rs:91-106 */
static void bit_rows(const float *c, int neg_rows, int neg_cols, uint16_t rows[16]) {
    float median = coefficient_median(c, neg_rows, neg_cols);
    for (int r = 0; r < 16; r++) {
        uint16_t bits = 0;
        for (int col = 0; col < 16; col++)
            if (apply_sign(c[r * 16 + col], r, col, neg_rows, neg_cols) > median) bits |= (uint16_t)(1u << col);
        rows[r] = bits;
    }
}

/* pdqhash.rs:140-151 */
static void transpose_bit_rows(const uint16_t in[16], uint16_t out[16]) {
    memset(out, 0, 32);
    for (int r = 0; r < 16; r++)
        for (int c = 0; c < 16; c++)
            if (in[r] & (1u << c)) out[c] |= (uint16_t)(1u << r);
}

/* pdqhash.rs:155-162 */
static void pack_bit_rows(const uint16_t rows[16], uint8_t *hash) {
    for (int r = 0; r < 16; r++) {
        hash[32 - 2 * r - 1] = (uint8_t)(rows[r] & 0xFF);
        hash[32 - 2 * r - 2] = (uint8_t)(rows[r] >> 8);
    }
}

/* pdqhash.rs:59-61 */
void orc_to_hash(const float *c, uint8_t *hash) {
    uint16_t rows[16];
    bit_rows(c, 0, 0, rows);
    pack_bit_rows(rows, hash);
}

/* pdqhash.rs:71-87 */
void orc_dihedral(const float *c, uint8_t *out) {
    uint16_t id[16], neg_cols[16], neg_rows[16], neg_both[16], t[16];
    bit_rows(c, 0, 0, id);
    bit_rows(c, 0, 1, neg_cols);
    bit_rows(c, 1, 0, neg_rows);
    bit_rows(c, 1, 1, neg_both);
    pack_bit_rows(id, out + 0 * 32);
    transpose_bit_rows(neg_rows, t);
    pack_bit_rows(t, out + 1 * 32);
    pack_bit_rows(neg_both, out + 2 * 32);
    transpose_bit_rows(neg_cols, t);
    pack_bit_rows(t, out + 3 * 32);
    pack_bit_rows(neg_cols, out + 4 * 32);
    pack_bit_rows(neg_rows, out + 5 * 32);
    transpose_bit_rows(id, t);
    pack_bit_rows(t, out + 6 * 32);
    transpose_bit_rows(neg_both, t);
    pack_bit_rows(t, out + 7 * 32);
}

/* pdqhash.rs:238-262 */
void orc_pdq_from_luma(const uint8_t *luma, uint32_t w, uint32_t h, float *coeffs, float *quality,
                       float *buf64_out) {
    size_t cols = w, rows = h;
    float *plane = (float *)scratch(SCR_PLANE, rows * cols * sizeof(float));
    for (size_t i = 0; i < rows * cols; i++) plane[i] = (float)luma[i]; /* :244 */
    size_t w_rows = (cols + BUF - 1) / BUF;                               /* :246 */
    size_t w_cols = (rows + BUF - 1) / BUF;                               /* :247 */
    orc_jarosz(plane, rows, cols, w_rows, w_cols, JAROSZ_PASSES);
    float buf64[BUF * BUF];
    orc_decimate64(plane, rows, cols, buf64);
    *quality = orc_quality(buf64, BUF, BUF);
    orc_dct64_to_16(buf64, coeffs);
    if (buf64_out) memcpy(buf64_out, buf64, sizeof(buf64));
}

/* pdqhash.rs:166-196 */
int orc_pdq_features(const uint8_t *px, int layout, uint32_t w, uint32_t h, float *coeffs,
                     float *quality, float *buf64) {
    if (w < MIN_HASHABLE_DIM || h < MIN_HASHABLE_DIM) return 1; /* None */
    size_t n = (size_t)w * h;
    uint8_t *luma_owned = NULL;
    const uint8_t *luma = px;
    if (layout != ORC_LAYOUT_LUMA8) { /* :172-175 */
        luma_owned = (uint8_t *)scratch(SCR_LUMA, n);
        orc_luma601(px, layout, n, luma_owned);
        luma = luma_owned;
    }
    uint8_t *resized = NULL;
    uint32_t pw = w, ph = h;
    if (w > DOWNSAMPLE_DIMS || h > DOWNSAMPLE_DIMS) { /* :181-188 */
        uint32_t nw, nh;
        orc_target_dimensions(w, h, DOWNSAMPLE_DIMS, &nw, &nh);
        resized = (uint8_t *)scratch(SCR_RESIZED, (size_t)nw * nh);
        if (orc_resize_box_u8(luma, (int)w, (int)h, resized, (int)nw, (int)nh) == 0) {
            luma = resized;
            pw = nw;
            ph = nh;
        } /* else: hash at full resolution */
    }
    orc_pdq_from_luma(luma, pw, ph, coeffs, quality, buf64);
    return 0;
}

/* scanner.rs:1416-1418: (q*100).round().clamp(0,100) as u16; f32::round is half away from zero */
uint16_t orc_quality_100(float q) {
    float v = roundf(q * 100.0f);
    if (v < 0.0f) v = 0.0f;
    if (v > 100.0f) v = 100.0f;
    return (uint16_t)v;
}

/* ---- multithreaded batch driver: one image per task (scanner.rs:1202-1205) ---- */
typedef struct {
    const uint8_t *px;
    int layout;
    size_t n;
    uint32_t w, h;
    size_t pitch;
    uint8_t *out_hash;
    float *out_quality;
    float *out_coeffs;
    uint8_t *out_dihedral;
    uint8_t *out_valid;
    size_t next;
    pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *job = (batch_job *)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        size_t i = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (i >= job->n) break;
        float coeffs[256], q = 0.0f;
        int none = orc_pdq_features(job->px + i * job->pitch, job->layout, job->w, job->h, coeffs, &q, NULL);
        if (job->out_valid) job->out_valid[i] = none ? 0 : 1;
        if (none) {
            memset(coeffs, 0, sizeof(coeffs));
            if (job->out_hash) memset(job->out_hash + i * 32, 0, 32);
            if (job->out_dihedral) memset(job->out_dihedral + i * 256, 0, 256);
        } else {
            if (job->out_hash) orc_to_hash(coeffs, job->out_hash + i * 32);
            if (job->out_dihedral) orc_dihedral(coeffs, job->out_dihedral + i * 256);
        }
        if (job->out_quality) job->out_quality[i] = q;
        if (job->out_coeffs) memcpy(job->out_coeffs + i * 256, coeffs, sizeof(coeffs));
    }
    scratch_release(); /* worker threads live for one batch */
    return NULL;
}

int orc_pdq_batch_mt(const uint8_t *px, int layout, size_t n, uint32_t w, uint32_t h, size_t img_pitch,
                     int threads, uint8_t *out_hash, float *out_quality, float *out_coeffs,
                     uint8_t *out_dihedral, uint8_t *out_valid) {
    pthread_once(&g_dct_once, dct_init);
    batch_job job = {px, layout, n, w, h, img_pitch, out_hash, out_quality, out_coeffs, out_dihedral, out_valid, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tids[256];
    for (int t = 0; t < threads; t++) pthread_create(&tids[t], NULL, batch_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);
    return 0;
}
