/*
 * rupphash_b200.h -- C ABI of librupphash_b200.so: the two data-parallel hot paths of
 * phdupes (Safari77/rupphash) as hand-written CUDA for sm_100a.
 *
 * The reference has no FFI for these paths (they are plain in-crate Rust functions), so
 * each entry point below names the Rust function(s) it replaces, file:line relative to
 * the reference tree.  INTEGRATION.md shows the `extern "C"` block + build.rs a
 * maintainer would add on the Rust side.
 *
 * Conventions
 *   - Every function returns an rh_status (0 = ok, < 0 = error); no exception or abort
 *     crosses the boundary.  rh_last_error(ctx) gives the message of the last failure.
 *   - The caller allocates and frees every in/out buffer.  Each pointer may be host
 *     (pageable or pinned) or device memory of the ctx's device; the library detects
 *     which (cudaPointerGetAttributes) and stages host buffers itself.  Inputs are not
 *     retained after return.  Calls are synchronous: results are in the caller's memory
 *     when the call returns (device outputs: the work is complete on the ctx stream).
 *   - A ctx is bound to one CUDA device and is NOT thread-safe (one in-flight call per
 *     ctx); several ctxs may coexist (one per decode thread pool / per GPU).
 *   - There is no CPU fallback: without a usable CUDA device rh_ctx_create fails.
 */
#ifndef RUPPHASH_B200_H
#define RUPPHASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rh_ctx rh_ctx;

typedef enum {
    RH_OK = 0,
    RH_EINVAL = -1,       /* bad argument (incl. threshold > 63: scanner.rs:1650-1655 assert) */
    RH_ECUDA = -2,        /* CUDA runtime error */
    RH_ENOMEM = -3,       /* allocation failed */
    RH_EUNSUPPORTED = -4, /* valid in the reference but not implemented on the device path */
    RH_ENCCL = -5         /* NCCL could not be loaded / initialised, or a collective failed */
} rh_status;

/* pixel layouts accepted by the hashers (pdqhash.rs:268-284: Rgb8 / Rgba8 / Luma8) */
typedef enum { RH_LAYOUT_RGB8 = 0, RH_LAYOUT_RGBA8 = 1, RH_LAYOUT_LUMA8 = 2 } rh_layout;

#define RH_PDQ_HASH_BYTES 32      /* pdqhash.rs:23 HASH_LENGTH */
#define RH_PDQ_COEFFS 256         /* pdqhash.rs:22 DCT_OUTPUT_MATRIX_SIZE */
#define RH_MAX_SIMILARITY_64 15   /* hamminghash.rs:5 */
#define RH_MAX_SIMILARITY_256 63  /* hamminghash.rs:8 */
#define RH_PDQ_MIN_QUALITY 50     /* scanner.rs:1579-1594 low-confidence rule (quality_100 < 50) */

/* ------------------------------------------------------------------ context ---- */

/* Create a context on CUDA device `device` (index as seen by this process). */
int rh_ctx_create(int device, rh_ctx **out);
int rh_ctx_destroy(rh_ctx *ctx);
/* Run all work of this ctx on `cuda_stream` (a cudaStream_t owned by the caller, e.g.
 * torch's current stream).  NULL restores the ctx-owned stream. */
int rh_ctx_set_stream(rh_ctx *ctx, void *cuda_stream);
int rh_ctx_sync(rh_ctx *ctx);
/* Tuning knobs for benchmarks and A/B runs (never needed for correct results; the defaults are the
 * product path): "hamming.prefilter" (-1 = chosen on the device from the sampled selectivity, 0..7 = pin the kernel variant),
 * "pdq.force_generic", "pdq.prefetch", "pdq.prefetch_rows", "pdq.phase_clocks", "pdq.variant". */
int rh_ctx_set_option(rh_ctx *ctx, const char *key, int value);
const char *rh_last_error(const rh_ctx *ctx);
const char *rh_version(void);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
uint64_t rh_kernel_launches(const rh_ctx *ctx);
/* Device time of the most recent call's dominant kernel(s), measured with CUDA events on
 * the launching stream: [0] = ms, [1] = units processed (images or pairs). */
int rh_last_kernel_time(const rh_ctx *ctx, double *ms, double *units);
/* The tile-kernel variant the most recent rh_hamming_group / _shard / _edges call of this ctx chose on the device
 * from its sampled selectivity (or the pinned one): 0 = full 256-bit distance for every pair (4 POPC), 3 / 4 = exact
 * 96- / 128-bit prefix (2 / 3 POPC), 2 / 1 = OR lower bound over the first 64 / 128 bits (1 POPC; strict thresholds),
 * 5 / 6 = OR lower bound over the first 160 / 192 bits (2 POPC), 7 = OR lower
 * bound over all 256 bits (3 POPC); for rh_hamming_group_u64: 1 = one-POPC OR bound, 0 = exact distance for every
 * pair; -1 before the first search.  Every variant gives identical results. */
int rh_hamming_last_variant(const rh_ctx *ctx);

/* pinned host staging the caller may use for the scanner-style batching pipeline */
int rh_alloc_pinned(size_t bytes, void **out);
int rh_free_pinned(void *p);

/* ---------------------------------------------------------------------- PDQ ---- */

/*
 * Batched pdqhash::generate_pdq_features + PdqFeatures::to_hash (+ optionally
 * generate_dihedral_hashes) -- pdqhash.rs:166-196, :199-201, :59-61, :71-87; the call
 * site it replaces is scanner.rs:1409-1412 (one call per image there, one call per
 * uniform-size batch here).
 *
 *   pixels      n images of w x h pixels, `layout` channels interleaved, rows row_pitch
 *               bytes apart (0 = tight), images img_pitch bytes apart (0 = tight)
 *   out_hash    n x 32 bytes, byte order of pack_bit_rows (pdqhash.rs:155-162)   [or NULL]
 *   out_quality n floats in [0,1] (pdqhash.rs:445-460)                            [or NULL]
 *   out_coeffs  n x 256 floats, PdqFeatures.coefficients row-major (pdqhash.rs:50) [or NULL]
 *   out_dihedral n x 8 x 32 bytes in the order of pdqhash.rs:77-86               [or NULL]
 *   out_valid   n bytes: 1 = Some, 0 = None (w < 5 or h < 5, pdqhash.rs:167-169)  [or NULL]
 *
 * Sizes: any w, h >= 5.  Images with a side above 512 px are pre-downsampled to the target
 * dimensions of pdqhash.rs:224-235 like the reference does: an exact 2x reduction in both axes
 * (1024x768 -> 512x384) is fused into the front end as two rounded halving passes; any other
 * ratio runs fast_image_resize's fixed-point Box convolution as restated by the oracle (that
 * crate is not in the reference tree: parity with it is unpinned).
 */
int rh_pdq_hash_batch(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h,
                      size_t row_pitch, size_t img_pitch, uint8_t *out_hash, float *out_quality,
                      float *out_coeffs, uint8_t *out_dihedral, uint8_t *out_valid);

/*
 * The same, queued: returns as soon as the copies and kernels are queued on the ctx stream.  The
 * caller's buffers (page-locked for real overlap) must stay valid and untouched until rh_ctx_wait
 * (ctx, *ticket) -- or rh_ctx_sync -- returns.  Calls on one ctx complete in order; up to 8 may be
 * outstanding.  With two batches in flight the host-to-device copy of batch k+1 runs under the kernels
 * of batch k: the scanner-style pipeline (scanner.rs:1202-1211: decode workers feed, one submitter
 * per ctx) is the intended caller.
 */
int rh_pdq_hash_batch_async(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h,
                            size_t row_pitch, size_t img_pitch, uint8_t *out_hash, float *out_quality,
                            float *out_coeffs, uint8_t *out_dihedral, uint8_t *out_valid,
                            uint64_t *ticket);
int rh_ctx_wait(rh_ctx *ctx, uint64_t ticket);

/* PdqFeatures::to_hash over n cached coefficient blocks (pdqhash.rs:59-61) */
int rh_pdq_hash_from_coeffs(rh_ctx *ctx, const float *coeffs, int64_t n, uint8_t *out_hash);
/* PdqFeatures::generate_dihedral_hashes over n blocks (pdqhash.rs:71-87; callers
 * scanner.rs:1622, :2223) */
int rh_pdq_dihedral_from_coeffs(rh_ctx *ctx, const float *coeffs, int64_t n, uint8_t *out_dihedral);
/* The 64x64 -> hash tail alone (quality + dct64_to_16 + to_hash), pdqhash.rs:258-260;
 * lets tests feed the reference's LCG buffers (pdqhash.rs:582-628). */
int rh_pdq_from_buffer64(rh_ctx *ctx, const float *buf64x64, int64_t n, uint8_t *out_hash,
                         float *out_quality, float *out_coeffs, uint8_t *out_dihedral);

/* -------------------------------------------------------------------- pHash ---- */

/* phash.rs:150-255: pure bit permutations, host-side, no ctx needed. */
uint64_t rh_phash_rotate_90(uint64_t h);
uint64_t rh_phash_rotate_180(uint64_t h);
uint64_t rh_phash_rotate_270(uint64_t h);
uint64_t rh_phash_flip_horizontal(uint64_t h);
void rh_phash_dihedral(uint64_t h, uint64_t out[8]);
uint64_t rh_phash_rotation_invariant(uint64_t h);
/* DctPhash::hash_image (phash.rs:48-83) over a batch; out_dihedral (n x 8) may be NULL.
 * PARITY UNVERIFIED against the real crates: the resize (image 0.25, Triangle), the Rec.709 luma and the DCT
 * (rustdct) are not in the reference tree; this path is bit-exact with the oracle's restatement of them only
 * (rustdct's butterfly summation order can differ in the last ulp, which can flip bits next to the median). */
int rh_phash_batch(rh_ctx *ctx, const uint8_t *pixels, int layout, int64_t n, int w, int h,
                   size_t row_pitch, size_t img_pitch, uint64_t *out_hash, uint64_t *out_dihedral);

/* ------------------------------------------------------------------ Hamming ---- */

/* HammingHash::hamming_distance for [u8;32] (hamminghash.rs:55-58) over n pairs
 * a[i] vs b[i]; callers scanner.rs:1885, :2228, :2236. */
int rh_hamming_distances(rh_ctx *ctx, const uint8_t *a, const uint8_t *b, int64_t n, uint32_t *out);
/* same for u64 hashes (hamminghash.rs:34-36) */
int rh_hamming_distances_u64(rh_ctx *ctx, const uint64_t *a, const uint64_t *b, int64_t n, uint32_t *out);

/*
 * scanner::group_files_generic (scanner.rs:1640-1817): edge phase + union-find, as an
 * exact all-pairs search (identical edge multiset to the reference's MIH probing for
 * every allowed similarity).
 *
 *   hashes      n x 32 bytes (file.pdqhash; rows without a hash are ignored)
 *   has_hash    n bytes or NULL (= all Some)                          scanner.rs:1658-1662
 *   variants    n x 8 x 32 bytes or NULL.  NULL: every file queries with its own hash
 *               (scanner.rs:1624-1627); else the n_variants[i] leading variants of file i
 *               (clamped to 1..8; 8 when n_variants is NULL) are its queries
 *               (scanner.rs:1615-1623: a file always queries with at least one hash).
 *   low_conf    n bytes or NULL: is_low_confidence (scanner.rs:1631-1636); a pair with a
 *               low-confidence side only matches at distance 0 (scanner.rs:1699,1721)
 *   similarity  <= 63, else RH_EINVAL
 *   out_label   n x u32: smallest file index of i's connected component (the canonical
 *               form of scanner.rs:1809-1817's unordered groups_map; groups = labels
 *               shared by > 1 file, members ascending)
 *   out_edge_count  edges.len() = comparison_count (scanner.rs:1778), per-variant
 *               duplicates included
 */
int rh_hamming_group(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash,
                     const uint8_t *variants, const uint8_t *n_variants, const uint8_t *low_conf,
                     int64_t n, uint32_t similarity, uint32_t *out_label, uint64_t *out_edge_count);

/*
 * One rank's share of the same search: the (query-block, candidate-block) tiles t with
 * owner(t) == rank out of `world`.  Writes the rank-local forest (out_parent[i] = smallest
 * index reachable from i through this rank's edges) and this rank's edge count.  The
 * ranks' forests are exchanged by the caller (NCCL all-gather of n x u32 per rank) and
 * combined with rh_uf_merge; the edge counts are summed (all-reduce).
 */
int rh_hamming_group_shard(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash,
                           const uint8_t *variants, const uint8_t *n_variants,
                           const uint8_t *low_conf, int64_t n, uint32_t similarity, int rank,
                           int world, uint32_t *out_parent, uint64_t *out_edge_count);
/* parents: world x n u32 forests -> out_label as in rh_hamming_group. */
int rh_uf_merge(rh_ctx *ctx, const uint32_t *parents, int world, int64_t n, uint32_t *out_label);

/* Debug / verification view of the same search: the edge list itself (file indices, i < j,
 * unordered, one entry per matching (variant, j)).  At most edges_cap pairs are written;
 * *out_edge_count is always the full count. */
int rh_hamming_edges(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash,
                     const uint8_t *variants, const uint8_t *n_variants, const uint8_t *low_conf,
                     int64_t n, uint32_t similarity, uint32_t *out_edges, size_t edges_cap,
                     uint64_t *out_edge_count);

/* u64 flavour (hashes n x u64; variants n x 8 x u64 or NULL; similarity <= 15 is what the
 * reference's MIH can serve, hamminghash.rs:5; the exact search here accepts <= 63). */
int rh_hamming_group_u64(rh_ctx *ctx, const uint64_t *hashes, const uint8_t *has_hash,
                         const uint64_t *variants, const uint8_t *n_variants,
                         const uint8_t *low_conf, int64_t n, uint32_t similarity,
                         uint32_t *out_label, uint64_t *out_edge_count);

/*
 * hamminghash::find_groups (hamminghash.rs:191-271): adjacency with d <= max_dist, then
 * the sequential greedy star clustering.  width_bits is 64 or 256.  The device finds the
 * exact adjacency (the reference's probing is complete only for max_dist <= 31 / 15;
 * above that this returns a superset).  Output is CSR: group g holds
 * members[group_offsets[g] .. group_offsets[g+1]), seed first, then ascending.
 * members_cap / groups_cap are the capacities of the caller's arrays (n and n/2+1 always
 * suffice); *n_groups receives the number of groups.  The adjacency is materialised on the host: inputs with
 * more than 2^30 edges (e.g. a million identical hashes) return RH_EUNSUPPORTED -- rh_hamming_group never
 * materialises edges and is the entry point for such libraries.
 */
int rh_find_groups(rh_ctx *ctx, const uint8_t *hashes, int64_t n, int width_bits, uint32_t max_dist,
                   uint32_t *members, size_t members_cap, uint32_t *group_offsets,
                   size_t groups_cap, size_t *n_groups);

/*
 * max_dist of analyze_group_with_features (scanner.rs:2217-2241) for many groups at once: for
 * group g, max over its members of (min over the pivot's variants of the Hamming distance).
 *   pivot_variants    n_groups x 8 x 32 bytes: generate_dihedral_hashes of each group's pivot
 *                     (or just its hash in slot 0 when the pivot has no cached coefficients,
 *                     scanner.rs:2231-2237)
 *   n_pivot_variants  n_groups bytes (1..8) or NULL (= 8)
 *   member_hashes     n_members x 32 bytes, member_group n_members x u32 (group of each member;
 *                     members without a hash are simply left out, scanner.rs:2221)
 *   out_max_dist      n_groups x u32 (0 for a group without members)
 * Choosing the pivot (the sort by duplicate status / stem, scanner.rs:2194-2214) is path and
 * metadata logic and stays on the host.
 */
int rh_group_max_dist(rh_ctx *ctx, const uint8_t *pivot_variants, const uint8_t *n_pivot_variants,
                      const uint8_t *member_hashes, const uint32_t *member_group, int64_t n_members,
                      int64_t n_groups, uint32_t *out_max_dist);

/* ------------------------------------------------------- several GPUs, one process ---- */

/*
 * rh_group: the GPUs of one box behind one handle, for a caller that is a single process (the
 * reference calls group_with_pdqhash once from its scan thread, scanner.rs:1550-1551, :1827-1832).
 * One rh_ctx per device, peer access between all of them, NCCL communicators from
 * ncclCommInitAll (NCCL is loaded with dlopen here: the library has no link-time NCCL dependency).
 *   devices   CUDA device indices, or NULL for devices 0 .. n_dev-1 (n_dev <= 0 with NULL: all)
 *   flags     RH_GROUP_NO_NCCL: exchange with cudaMemcpyPeerAsync instead of NCCL collectives
 *             RH_GROUP_STEAL_TILES: the GPUs claim tiles from one counter in the first GPU's memory with
 *             NVLink atomics (work stealing; absorbs unequal GPUs).  Default: tile t belongs to GPU
 *             t mod n_dev, which measured 1 % faster on 8 equal B200s (RH_GROUP_STATIC_TILES = the default)
 * Errors: RH_ENCCL when NCCL cannot be loaded / initialised; rh_group_last_error for the message.
 * A group is not thread-safe (one in-flight call); rh_group_ctx gives the per-device contexts for
 * single-device calls between group calls.
 */
typedef struct rh_group rh_group;
#define RH_GROUP_NO_NCCL 1u
#define RH_GROUP_STATIC_TILES 2u
#define RH_GROUP_STEAL_TILES 4u
int rh_group_create(const int *devices, int n_dev, unsigned flags, rh_group **out);
int rh_group_destroy(rh_group *g);
int rh_group_size(const rh_group *g);
rh_ctx *rh_group_ctx(rh_group *g, int i);
const char *rh_group_last_error(const rh_group *g);
/* NCCL version in use (0 = peer copies) and whether tiles are claimed from one pool */
int rh_group_info(const rh_group *g, int *nccl_version, int *work_stealing);
/* Times of the last group call in ms: [0] wall time of rh_hamming_group_multi, [1] / [2] tile kernel
 * on the slowest / fastest GPU (CUDA events), [3] sum of the tile-kernel times over the GPUs,
 * [4] wall time of rh_pdq_hash_batch_multi; the first GPU's timeline of the last group call (CUDA events):
 * [5] inputs (copies + replication), [6] dense arrays, [7] exchange of forests / counts after its tile kernel,
 * [8] merge + copy-out, [9] first to last mark. */
int rh_group_last_times(const rh_group *g, double *out, int n_out);

/*
 * scanner::group_files_generic (scanner.rs:1640-1817) over every GPU of the group; same arguments,
 * same results (labels, comparison_count) as rh_hamming_group, bit-identical for any number of GPUs.
 * Inputs: host memory (each GPU copies one slice over its own PCIe link, an NCCL all-gather over
 * NVLink replicates them) or memory of one GPU of the group (broadcast from there).  The N x N pair
 * matrix is tiled across the GPUs, the rank-local forests are all-gathered (n x u32 per GPU), the
 * edge counts all-reduced, the forests merged on the first GPU.  out_label: host memory or memory
 * of the group's first GPU.
 */
int rh_hamming_group_multi(rh_group *g, const uint8_t *hashes, const uint8_t *has_hash,
                           const uint8_t *variants, const uint8_t *n_variants, const uint8_t *low_conf,
                           int64_t n, uint32_t similarity, uint32_t *out_label, uint64_t *out_edge_count);

/* rh_pdq_hash_batch with the batch cut into one contiguous slice per GPU (host buffers; images are
 * independent, so there is no exchange: scanner.rs:1202-1205's par_iter over files). */
int rh_pdq_hash_batch_multi(rh_group *g, const uint8_t *pixels, int layout, int64_t n, int w, int h,
                            size_t row_pitch, size_t img_pitch, uint8_t *out_hash, float *out_quality,
                            float *out_coeffs, uint8_t *out_dihedral, uint8_t *out_valid);

/* ------------------------------------------------------------- measurement ---- */

/* Integer-pipe and copy peaks measured on this device, used as roofline denominators
 * where MEASURED_PEAKS.json has none: out[0] = POPC.32 lane-ops/s, out[1] = LOP3 lane-ops/s,
 * out[2] = pinned H2D GB/s, out[3] = device copy GB/s (read+write bytes). */
int rh_measure_peaks(rh_ctx *ctx, double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* RUPPHASH_B200_H */
