/*
 * oracle/oracle_group.c -- CPU restatement of /root/reference/src/hamminghash.rs and
 * the grouping core of /root/reference/src/scanner.rs:1640-1817.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* HammingHash trait (hamminghash.rs:11-63)                                  */

/* hamminghash.rs:55-58 */
uint32_t orc_hamming256(const uint8_t *a, const uint8_t *b) {
    uint32_t d = 0;
    for (int i = 0; i < 32; i++) d += (uint32_t)__builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}
/* hamminghash.rs:34-36 */
uint32_t orc_hamming64(uint64_t a, uint64_t b) { return (uint32_t)__builtin_popcountll(a ^ b); }

typedef struct {
    int width_bits;   /* 64 | 256 */
    int num_chunks;   /* 8 | 16         hamminghash.rs:24,45 */
    int num_buckets;  /* 256 | 65536    hamminghash.rs:25,46 */
    int chunk_bits;   /* 8 | 16         hamminghash.rs:38-40,60-62 */
    int bytes;
} hh_traits;

static hh_traits traits_for(int width_bits) {
    hh_traits t;
    t.width_bits = width_bits;
    if (width_bits == 64) {
        t.num_chunks = 8; t.num_buckets = 256; t.chunk_bits = 8; t.bytes = 8;
    } else {
        t.num_chunks = 16; t.num_buckets = 65536; t.chunk_bits = 16; t.bytes = 32;
    }
    return t;
}

/* hamminghash.rs:29-31 (u64: byte k of the integer) and :50-53 ([u8;32]: LE u16 of bytes 2k,2k+1).
 * A little-endian u64 in memory has byte k at offset k, so both are plain byte reads. */
static inline uint16_t get_chunk(const hh_traits *t, const uint8_t *h, int k) {
    if (t->width_bits == 64) return h[k];
    return (uint16_t)(h[2 * k] | ((uint16_t)h[2 * k + 1] << 8));
}

static inline uint32_t hh_distance(const hh_traits *t, const uint8_t *a, const uint8_t *b) {
    if (t->width_bits == 64) {
        uint64_t x, y;
        memcpy(&x, a, 8);
        memcpy(&y, b, 8);
        return orc_hamming64(x, y);
    }
    /* same value as orc_hamming256, word-at-a-time so the CPU baseline is not handicapped */
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return (uint32_t)(__builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
                      __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]));
}

/* ------------------------------------------------------------------------- */
/* MIHIndex (hamminghash.rs:82-149)                                          */

struct orc_mih {
    hh_traits t;
    size_t n;
    uint8_t *db_hashes;
    uint32_t *offsets; /* num_chunks*num_buckets + 1 */
    uint32_t *values;  /* n * num_chunks */
    size_t n_offsets;
};

orc_mih *orc_mih_new(const uint8_t *hashes, size_t n, int width_bits) {
    if (width_bits != 64 && width_bits != 256) return NULL;
    orc_mih *m = (orc_mih *)calloc(1, sizeof(*m));
    m->t = traits_for(width_bits);
    m->n = n;
    size_t nb = (size_t)m->t.num_chunks * m->t.num_buckets;
    m->n_offsets = nb + 1;
    m->offsets = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
    m->db_hashes = (uint8_t *)malloc(n * m->t.bytes + 1);
    memcpy(m->db_hashes, hashes, n * m->t.bytes);
    /* count phase :95-103 */
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < m->t.num_chunks; k++) {
            size_t flat = (size_t)k * m->t.num_buckets + get_chunk(&m->t, hashes + i * m->t.bytes, k);
            m->offsets[flat + 1] += 1;
        }
    /* prefix sum :106-108 */
    for (size_t i = 1; i < nb + 1; i++) m->offsets[i] += m->offsets[i - 1];
    m->values = (uint32_t *)calloc((size_t)m->offsets[nb] + 1, sizeof(uint32_t));
    /* fill phase :113-123 */
    uint32_t *cursor = (uint32_t *)malloc((nb + 1) * sizeof(uint32_t));
    memcpy(cursor, m->offsets, (nb + 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < m->t.num_chunks; k++) {
            size_t flat = (size_t)k * m->t.num_buckets + get_chunk(&m->t, hashes + i * m->t.bytes, k);
            m->values[cursor[flat]++] = (uint32_t)i;
        }
    free(cursor);
    return m;
}

void orc_mih_free(orc_mih *m) {
    if (!m) return;
    free(m->db_hashes);
    free(m->offsets);
    free(m->values);
    free(m);
}

/* hamminghash.rs:133-138 */
const uint32_t *orc_mih_bucket(const orc_mih *m, int chunk, uint16_t value, size_t *len) {
    size_t flat = (size_t)chunk * m->t.num_buckets + value;
    *len = m->offsets[flat + 1] - m->offsets[flat];
    return m->values + m->offsets[flat];
}
const uint32_t *orc_mih_offsets(const orc_mih *m, size_t *len) {
    *len = m->n_offsets;
    return m->offsets;
}
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* SparseBitSet (hamminghash.rs:152-189)                                     */

typedef struct {
    uint64_t *data;
    size_t *dirty;
    size_t n_dirty, cap_dirty;
} sparse_bitset;

static void sbs_init(sparse_bitset *s, size_t size) {
    s->data = (uint64_t *)calloc((size + 63) / 64 + 1, sizeof(uint64_t));
    s->cap_dirty = 512;
    s->dirty = (size_t *)malloc(s->cap_dirty * sizeof(size_t));
    s->n_dirty = 0;
}
static void sbs_free(sparse_bitset *s) {
    free(s->data);
    free(s->dirty);
}
/* returns previous state (:163-180) */
static inline int sbs_set(sparse_bitset *s, size_t idx) {
    size_t w = idx / 64;
    uint64_t mask = 1ull << (idx % 64);
    uint64_t word = s->data[w];
    int was = (word & mask) != 0;
    if (!was) {
        if (word == 0) {
            if (s->n_dirty == s->cap_dirty) {
                s->cap_dirty *= 2;
                s->dirty = (size_t *)realloc(s->dirty, s->cap_dirty * sizeof(size_t));
            }
            s->dirty[s->n_dirty++] = w;
        }
        s->data[w] = word | mask;
    }
    return was;
}
static inline void sbs_clear(sparse_bitset *s) {
    for (size_t i = 0; i < s->n_dirty; i++) s->data[s->dirty[i]] = 0;
    s->n_dirty = 0;
}

typedef struct {
    uint32_t *v;
    size_t n, cap;
} u32vec;
static inline void vec_push(u32vec *v, uint32_t x) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->v = (uint32_t *)realloc(v->v, v->cap * sizeof(uint32_t));
    }
    v->v[v->n++] = x;
}

/* ------------------------------------------------------------------------- */
/* find_groups (hamminghash.rs:191-271)                                      */

typedef struct {
    const orc_mih *m;
    uint32_t max_dist;
    u32vec *adj;
    size_t next;
    pthread_mutex_t mu;
} fg_job;

static void fg_check_bucket(const orc_mih *m, int k, uint16_t val, size_t i, const uint8_t *q,
                            uint32_t max_dist, sparse_bitset *visited, u32vec *results) {
    size_t len;
    const uint32_t *bucket = orc_mih_bucket(m, k, val, &len);
    for (size_t b = 0; b < len; b++) {
        size_t d = bucket[b];
        if (d == i) continue;               /* :216-218 */
        if (sbs_set(visited, d)) continue;  /* :220-222 */
        if (hh_distance(&m->t, q, m->db_hashes + d * m->t.bytes) <= max_dist) vec_push(results, (uint32_t)d);
    }
}

static void *fg_worker(void *arg) {
    fg_job *job = (fg_job *)arg;
    const orc_mih *m = job->m;
    sparse_bitset visited;
    sbs_init(&visited, m->n);
    uint32_t chunk_tolerance = job->max_dist / (uint32_t)m->t.num_chunks; /* :193 */
    for (;;) {
        pthread_mutex_lock(&job->mu);
        size_t lo = job->next;
        job->next += 1024;
        pthread_mutex_unlock(&job->mu);
        if (lo >= m->n) break;
        size_t hi = lo + 1024 < m->n ? lo + 1024 : m->n;
        for (size_t i = lo; i < hi; i++) {
            sbs_clear(&visited);
            const uint8_t *q = m->db_hashes + i * m->t.bytes;
            for (int k = 0; k < m->t.num_chunks; k++) {
                uint16_t qc = get_chunk(&m->t, q, k);
                fg_check_bucket(m, k, qc, i, q, job->max_dist, &visited, &job->adj[i]);
                if (chunk_tolerance >= 1) /* :233-237 */
                    for (int bit = 0; bit < m->t.chunk_bits; bit++)
                        fg_check_bucket(m, k, (uint16_t)(qc ^ (1u << bit)), i, q, job->max_dist, &visited, &job->adj[i]);
            }
        }
    }
    sbs_free(&visited);
    return NULL;
}

int orc_find_groups(const orc_mih *m, uint32_t max_dist, int threads, uint32_t **members_out,
                    uint32_t **offsets_out, size_t *n_groups_out) {
    size_t n = m->n;
    u32vec *adj = (u32vec *)calloc(n + 1, sizeof(u32vec));
    fg_job job = {m, max_dist, adj, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tids[256];
    for (int t = 0; t < threads; t++) pthread_create(&tids[t], NULL, fg_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);

    /* greedy clustering :245-270 */
    uint8_t *visited = (uint8_t *)calloc(n + 1, 1);
    u32vec members = {0}, offsets = {0};
    vec_push(&offsets, 0);
    for (size_t i = 0; i < n; i++) {
        if (visited[i] || adj[i].n == 0) continue;
        size_t start = members.n;
        vec_push(&members, (uint32_t)i);
        visited[i] = 1;
        for (size_t a = 0; a < adj[i].n; a++) {
            uint32_t nb = adj[i].v[a];
            if (!visited[nb]) {
                visited[nb] = 1;
                vec_push(&members, nb);
            }
        }
        if (members.n - start > 1) vec_push(&offsets, (uint32_t)members.n);
        else members.n = start;
    }
    for (size_t i = 0; i < n; i++) free(adj[i].v);
    free(adj);
    free(visited);
    if (!members.v) members.v = (uint32_t *)malloc(4);
    *members_out = members.v;
    *offsets_out = offsets.v;
    *n_groups_out = offsets.n - 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* group_files_generic (scanner.rs:1640-1817)                                */

#define CHUNK_SIZE 2000 /* scanner.rs:1676 */

typedef struct {
    const uint8_t *hashes, *has_hash, *variants, *n_variants, *low_conf;
    size_t n;
    uint32_t similarity;
    int use_mih;
    const orc_mih *mih;
    const uint32_t *dense_to_sparse;
    u32vec *chunk_edges; /* per 2000-file chunk, (u,v) interleaved */
    size_t n_chunks;
    size_t next;
    size_t chunk_stride; /* 1 = every 2000-file chunk (the reference); k > 1: a bounded sample, every k-th chunk */
    size_t sample_files; /* 0 = whole chunks; m > 0: only the first m files of each searched chunk (timing sample) */
    pthread_mutex_t mu;
} grp_job;

static inline int file_has_hash(const grp_job *j, size_t i) { return !j->has_hash || j->has_hash[i]; }

static inline void grp_check_bucket(const grp_job *j, int k, uint16_t val, size_t i, const uint8_t *variant,
                                    uint32_t base_limit, sparse_bitset *visited, u32vec *edges) {
    size_t len;
    const uint32_t *bucket = orc_mih_bucket(j->mih, k, val, &len);
    for (size_t b = 0; b < len; b++) {
        uint32_t dense = bucket[b];
        size_t cand = j->dense_to_sparse[dense];              /* :1714 */
        if (cand <= i || sbs_set(visited, cand)) continue;    /* :1716-1718 */
        uint32_t limit = (j->low_conf && j->low_conf[cand]) ? 0 : base_limit; /* :1721 */
        if (hh_distance(&j->mih->t, variant, j->mih->db_hashes + (size_t)dense * 32) <= limit) { /* :1722 */
            vec_push(edges, (uint32_t)i);
            vec_push(edges, (uint32_t)cand);
        }
    }
}

static void *grp_worker(void *arg) {
    grp_job *j = (grp_job *)arg;
    sparse_bitset visited;
    sbs_init(&visited, j->n);
    uint8_t variants_buf[8 * 32];
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t chunk = j->next;
        j->next += j->chunk_stride ? j->chunk_stride : 1;
        pthread_mutex_unlock(&j->mu);
        if (chunk >= j->n_chunks) break;
        u32vec *edges = &j->chunk_edges[chunk];
        size_t lo = chunk * CHUNK_SIZE, hi = lo + CHUNK_SIZE < j->n ? lo + CHUNK_SIZE : j->n;
        if (j->sample_files && lo + j->sample_files < hi) hi = lo + j->sample_files;
        for (size_t i = lo; i < hi; i++) {
            if (!file_has_hash(j, i)) continue; /* :1690 */
            int count;                          /* :1692-1693, :1615-1628 */
            if (j->variants) {
                count = j->n_variants ? j->n_variants[i] : 8;
                if (count < 1) count = 1;
                if (count > 8) count = 8;
                memcpy(variants_buf, j->variants + i * 256, (size_t)count * 32);
            } else {
                count = 1;
                memcpy(variants_buf, j->hashes + i * 32, 32);
            }
            uint32_t base_limit = (j->low_conf && j->low_conf[i]) ? 0 : j->similarity; /* :1699 */
            for (int v = 0; v < count; v++) {
                const uint8_t *variant = variants_buf + v * 32;
                if (!j->use_mih) { /* independent brute-force statement of the same edge rule */
                    for (size_t cand = i + 1; cand < j->n; cand++) {
                        if (!file_has_hash(j, cand)) continue;
                        uint32_t limit = (j->low_conf && j->low_conf[cand]) ? 0 : base_limit;
                        if (orc_hamming256(variant, j->hashes + cand * 32) <= limit) {
                            vec_push(edges, (uint32_t)i);
                            vec_push(edges, (uint32_t)cand);
                        }
                    }
                    continue;
                }
                sbs_clear(&visited); /* :1702 */
                for (int k = 0; k < 16; k++) {
                    uint16_t q = (uint16_t)(variant[2 * k] | ((uint16_t)variant[2 * k + 1] << 8));
                    const int bits = 16;
                    grp_check_bucket(j, k, q, i, variant, base_limit, &visited, edges); /* R=0 :1729 */
                    if (j->similarity >= 16) /* R=1 :1732-1736 */
                        for (int a = 0; a < bits; a++)
                            grp_check_bucket(j, k, (uint16_t)(q ^ (1u << a)), i, variant, base_limit, &visited, edges);
                    if (j->similarity >= 32) /* R=2 :1739-1749 */
                        for (int a = 0; a < bits; a++)
                            for (int b = a + 1; b < bits; b++)
                                grp_check_bucket(j, k, (uint16_t)(q ^ (1u << a) ^ (1u << b)), i, variant, base_limit, &visited, edges);
                    if (j->similarity >= 48) /* R=3 :1752-1767 */
                        for (int a = 0; a < bits; a++)
                            for (int b = a + 1; b < bits; b++)
                                for (int c = b + 1; c < bits; c++)
                                    grp_check_bucket(j, k, (uint16_t)(q ^ (1u << a) ^ (1u << b) ^ (1u << c)), i, variant, base_limit, &visited, edges);
                }
            }
        }
    }
    sbs_free(&visited);
    return NULL;
}

/* scanner.rs:1783-1803 */
static size_t uf_find(size_t *parent, size_t i) {
    size_t root = i;
    while (root != parent[root]) root = parent[root];
    size_t curr = i;
    while (curr != root) {
        size_t next = parent[curr];
        parent[curr] = root;
        curr = next;
    }
    return root;
}
static void uf_union(size_t *parent, size_t i, size_t j) {
    size_t ri = uf_find(parent, i), rj = uf_find(parent, j);
    if (ri != rj) parent[ri] = rj;
}

/* canonical labels: min member of each component (SURVEY 8a "canonical form") */
static void labels_from_parent(size_t *parent, size_t n, uint32_t *out_label) {
    uint32_t *min_of_root = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) min_of_root[i] = UINT32_MAX;
    for (size_t i = 0; i < n; i++) {
        size_t r = uf_find(parent, i);
        if (min_of_root[r] == UINT32_MAX) min_of_root[r] = (uint32_t)i; /* i ascending => first = min */
    }
    for (size_t i = 0; i < n; i++) out_label[i] = min_of_root[uf_find(parent, i)];
    free(min_of_root);
}

static int group_generic_impl(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                              const uint8_t *n_variants, const uint8_t *low_conf, size_t n, uint32_t similarity,
                              int threads, int use_mih, uint32_t *out_label, uint64_t *out_edge_count,
                              uint32_t *edges_out, size_t edges_cap, size_t chunk_stride, size_t sample_files);

int orc_group_generic(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                      const uint8_t *n_variants, const uint8_t *low_conf, size_t n, uint32_t similarity,
                      int threads, int use_mih, uint32_t *out_label, uint64_t *out_edge_count,
                      uint32_t *edges_out, size_t edges_cap) {
    return group_generic_impl(hashes, has_hash, variants, n_variants, low_conf, n, similarity, threads, use_mih,
                              out_label, out_edge_count, edges_out, edges_cap, 1, 0);
}

/* Timing aid for bench.py's cpu_baseline on inputs whose full CPU search takes minutes: the same index
 * build and probe loop, but only every chunk_stride-th 2000-file chunk of query files is searched, and of each
 * of those only its first sample_files files (0 = all): many small work units, so that the sample keeps every
 * thread busy (labels then describe that sample's edges only; the caller scales the probe time). */
int orc_group_generic_sampled(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                              const uint8_t *n_variants, const uint8_t *low_conf, size_t n, uint32_t similarity,
                              int threads, size_t chunk_stride, size_t sample_files, uint32_t *out_label,
                              uint64_t *out_edge_count) {
    return group_generic_impl(hashes, has_hash, variants, n_variants, low_conf, n, similarity, threads, 1, out_label,
                              out_edge_count, NULL, 0, chunk_stride < 1 ? 1 : chunk_stride, sample_files);
}

static int group_generic_impl(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                              const uint8_t *n_variants, const uint8_t *low_conf, size_t n, uint32_t similarity,
                              int threads, int use_mih, uint32_t *out_label, uint64_t *out_edge_count,
                              uint32_t *edges_out, size_t edges_cap, size_t chunk_stride, size_t sample_files) {
    if (similarity > 63) return -1; /* assert scanner.rs:1650-1655 */
    for (size_t i = 0; i < n; i++) out_label[i] = (uint32_t)i;
    *out_edge_count = 0;

    /* valid entries :1658-1669 */
    size_t n_valid = 0;
    for (size_t i = 0; i < n; i++)
        if (!has_hash || has_hash[i]) n_valid++;
    if (n_valid == 0) return 0; /* :1664-1666 */
    uint8_t *dense_hashes = (uint8_t *)malloc(n_valid * 32);
    uint32_t *dense_to_sparse = (uint32_t *)malloc(n_valid * sizeof(uint32_t));
    size_t d = 0;
    for (size_t i = 0; i < n; i++)
        if (!has_hash || has_hash[i]) {
            memcpy(dense_hashes + d * 32, hashes + i * 32, 32);
            dense_to_sparse[d++] = (uint32_t)i;
        }

    grp_job job;
    memset(&job, 0, sizeof(job));
    job.hashes = hashes; job.has_hash = has_hash; job.variants = variants; job.n_variants = n_variants;
    job.low_conf = low_conf; job.n = n; job.similarity = similarity; job.use_mih = use_mih;
    job.mih = use_mih ? orc_mih_new(dense_hashes, n_valid, 256) : NULL; /* :1673 */
    job.dense_to_sparse = dense_to_sparse;
    job.n_chunks = (n + CHUNK_SIZE - 1) / CHUNK_SIZE;
    job.chunk_stride = chunk_stride;
    job.sample_files = sample_files;
    job.chunk_edges = (u32vec *)calloc(job.n_chunks + 1, sizeof(u32vec));
    pthread_mutex_init(&job.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tids[256];
    for (int t = 0; t < threads; t++) pthread_create(&tids[t], NULL, grp_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);

    /* flatten in chunk order (:1775-1776), count (:1778), sequential union-find (:1781-1807) */
    size_t *parent = (size_t *)malloc((n + 1) * sizeof(size_t));
    for (size_t i = 0; i < n; i++) parent[i] = i;
    uint64_t count = 0;
    for (size_t c = 0; c < job.n_chunks; c++) {
        u32vec *e = &job.chunk_edges[c];
        for (size_t k = 0; k + 1 < e->n; k += 2) {
            if (edges_out && count < edges_cap) {
                edges_out[2 * count] = e->v[k];
                edges_out[2 * count + 1] = e->v[k + 1];
            }
            count++;
            uf_union(parent, e->v[k], e->v[k + 1]);
        }
        free(e->v);
    }
    *out_edge_count = count;
    labels_from_parent(parent, n, out_label);

    free(parent);
    free(job.chunk_edges);
    if (job.mih) orc_mih_free((orc_mih *)job.mih);
    free(dense_hashes);
    free(dense_to_sparse);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Multi-rank emulation for the CPU (gloo) tests.                            */
/*
 * The pair matrix is cut into tiles of `tile` query-variants x `tile` candidate
 * files.  Query-variants are the flattened (file asc, variant asc) list of files
 * that have a hash.  A tile (qb, cb) is VALID when its largest candidate index
 * exceeds the file index of its first query (otherwise no j > i pair can exist).
 * Valid tiles are numbered row-major; tile t belongs to rank t % world.
 * (The device search owns tiles the same way when its ranks are static: tile t -> rank t % world.)
 */
int orc_group_tiles_rank(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                         const uint8_t *n_variants, const uint8_t *low_conf, size_t n, uint32_t similarity,
                         uint32_t tile, int rank, int world, uint32_t *out_parent, uint64_t *out_edge_count) {
    if (similarity > 63 || tile == 0 || world < 1) return -1;
    size_t *parent = (size_t *)malloc((n + 1) * sizeof(size_t));
    for (size_t i = 0; i < n; i++) parent[i] = i;
    /* flatten queries */
    size_t nq = 0;
    for (size_t i = 0; i < n; i++)
        if (!has_hash || has_hash[i]) nq += variants ? (n_variants ? (n_variants[i] < 1 ? 1 : (n_variants[i] > 8 ? 8 : n_variants[i])) : 8) : 1;
    uint32_t *qfile = (uint32_t *)malloc((nq + 1) * sizeof(uint32_t));
    const uint8_t **qhash = (const uint8_t **)malloc((nq + 1) * sizeof(uint8_t *));
    size_t q = 0;
    for (size_t i = 0; i < n; i++) {
        if (has_hash && !has_hash[i]) continue;
        int cnt = variants ? (n_variants ? (n_variants[i] < 1 ? 1 : (n_variants[i] > 8 ? 8 : n_variants[i])) : 8) : 1;
        for (int v = 0; v < cnt; v++) {
            qfile[q] = (uint32_t)i;
            qhash[q] = variants ? variants + i * 256 + (size_t)v * 32 : hashes + i * 32;
            q++;
        }
    }
    size_t n_qb = (nq + tile - 1) / tile, n_cb = (n + tile - 1) / tile;
    uint64_t t = 0, count = 0;
    for (size_t qb = 0; qb < n_qb; qb++) {
        size_t q0 = qb * tile, q1 = q0 + tile < nq ? q0 + tile : nq;
        for (size_t cb = 0; cb < n_cb; cb++) {
            size_t c0 = cb * tile, c1 = c0 + tile < n ? c0 + tile : n;
            if (!(c1 - 1 > qfile[q0])) continue; /* not valid */
            int mine = (int)(t % (uint64_t)world) == rank;
            t++;
            if (!mine) continue;
            for (size_t qq = q0; qq < q1; qq++) {
                size_t i = qfile[qq];
                uint32_t base_limit = (low_conf && low_conf[i]) ? 0 : similarity;
                for (size_t cand = c0 > i + 1 ? c0 : i + 1; cand < c1; cand++) {
                    if (has_hash && !has_hash[cand]) continue;
                    uint32_t limit = (low_conf && low_conf[cand]) ? 0 : base_limit;
                    if (orc_hamming256(qhash[qq], hashes + cand * 32) <= limit) {
                        count++;
                        /* min-root union so the forest matches what a GPU rank produces in spirit */
                        size_t ra = uf_find(parent, i), rb = uf_find(parent, cand);
                        if (ra != rb) {
                            if (ra < rb) parent[rb] = ra; else parent[ra] = rb;
                        }
                    }
                }
            }
        }
    }
    for (size_t i = 0; i < n; i++) out_parent[i] = (uint32_t)uf_find(parent, i);
    *out_edge_count = count;
    free(parent);
    free(qfile);
    free(qhash);
    return 0;
}

void orc_merge_parents(const uint32_t *parents, int world, size_t n, uint32_t *out_label) {
    size_t *parent = (size_t *)malloc((n + 1) * sizeof(size_t));
    for (size_t i = 0; i < n; i++) parent[i] = i;
    for (int g = 0; g < world; g++)
        for (size_t i = 0; i < n; i++) uf_union(parent, i, parents[(size_t)g * n + i]);
    labels_from_parent(parent, n, out_label);
    free(parent);
}
