/*
 * oracle/oracle.h -- CPU restatement ("oracle") of the two phdupes hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (rupphash_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or the
 * reported CPU baseline -- never as the thing shipped.
 *
 * What it restates (all file:line relative to /root/reference):
 *   src/pdqhash.rs:17-36,48-162,166-262,268-460   PDQ hashing
 *   src/phash.rs:48-83,137-255                    64-bit pHash (+ bit-level dihedral ops)
 *   src/hamminghash.rs:5-271                      HammingHash, MIHIndex, SparseBitSet, find_groups
 *   src/scanner.rs:1588-1594,1640-1817            low-confidence rule, group_files_generic
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - The Rust reference cannot be compiled here (no cargo/rustc) => there is no
 *     oracle/_ref build.
 *   - The reference ships NO absolute golden hash for any image (the tests/ .txt files are
 *     licence notes).  The oracle is pinned by every PORTABLE unit test the
 *     reference holds for these paths (relationship tests + KATs, tests/test_oracle_*.py)
 *     and by an independent numpy twin (oracle/np_twin.py).
 *   - PARITY UNPINNED for: absolute PDQ/pHash values on image files, the
 *     fast_image_resize pre-downsample for non-power-of-two ratios, and the whole
 *     pHash image path (image/rustdct crates are not vendored).
 *
 * Numerics rules: f32 everywhere, no FMA contraction (-ffp-contract=off), the
 * accumulation orders exactly as the reference writes them.
 */
#ifndef RUPPHASH_ORACLE_H
#define RUPPHASH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_LAYOUT_RGB8 = 0, ORC_LAYOUT_RGBA8 = 1, ORC_LAYOUT_LUMA8 = 2 };

/* ---------------------------------------------------------------- PDQ ---- */

/* pdqhash.rs:224-235 */
void orc_target_dimensions(uint32_t w, uint32_t h, uint32_t max_dim, uint32_t *nw, uint32_t *nh);
/* pdqhash.rs:268-284 */
void orc_luma601(const uint8_t *px, int layout, size_t n_pixels, uint8_t *out);
/* pdqhash.rs:203-220 -> fast_image_resize 6.1.0 Convolution(Box) on U8 (recalled, UNVERIFIED) */
int orc_resize_box_u8(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh);
/* pdqhash.rs:341-396 */
void orc_box_one_d(const float *in, size_t in_start, float *out, size_t out_start, size_t len,
                   size_t stride, size_t win);
/* pdqhash.rs:410-426 */
void orc_jarosz(float *buf, size_t rows, size_t cols, size_t w_rows, size_t w_cols, size_t nreps);
/* pdqhash.rs:428-443 */
void orc_decimate64(const float *in, size_t rows, size_t cols, float *out64x64);
/* pdqhash.rs:445-460 (generic R x C so the reference's 2x2 KAT can be run) */
float orc_quality(const float *buf, size_t R, size_t C);
/* pdqhash.rs:287-304 */
void orc_dct_matrix(float *D16x64);
/* pdqhash.rs:306-336 */
void orc_dct64_to_16(const float *in64x64, float *out256);
/* pdqhash.rs:59-61,91-124,155-162 */
void orc_to_hash(const float *coeffs256, uint8_t *hash32);
/* pdqhash.rs:71-87 */
void orc_dihedral(const float *coeffs256, uint8_t *out8x32);
/* pdqhash.rs:238-262.  buf64 may be NULL. */
void orc_pdq_from_luma(const uint8_t *luma, uint32_t w, uint32_t h, float *coeffs256,
                       float *quality, float *buf64);
/* pdqhash.rs:166-196.  returns 0 = Some, 1 = None (too small).  luma_out/lw/lh optional. */
int orc_pdq_features(const uint8_t *px, int layout, uint32_t w, uint32_t h, float *coeffs256,
                     float *quality, float *buf64);
/* scanner.rs:1416-1418 */
uint16_t orc_quality_100(float q);

/* One image per task over `threads` pthreads (mirrors par_iter, scanner.rs:1203).
 * Any of out_hash / out_quality / out_coeffs / out_dihedral may be NULL. */
int orc_pdq_batch_mt(const uint8_t *px, int layout, size_t n, uint32_t w, uint32_t h,
                     size_t img_pitch, int threads, uint8_t *out_hash, float *out_quality,
                     float *out_coeffs, uint8_t *out_dihedral, uint8_t *out_valid);

/* -------------------------------------------------------------- pHash ---- */

/* phash.rs:150-255 (bit permutations, exactly portable) */
uint64_t orc_phash_rot90(uint64_t h);
uint64_t orc_phash_rot180(uint64_t h);
uint64_t orc_phash_rot270(uint64_t h);
uint64_t orc_phash_flip_h(uint64_t h);
void orc_phash_dihedral(uint64_t h, uint64_t *out8);
uint64_t orc_phash_rot_invariant(uint64_t h);
/* phash.rs:55-83 on an already resized 32x32 luma plane (naive DCT-II order; the
 * rustdct butterfly order is not reproduced => last-ulp differences possible). */
uint64_t orc_phash_from_luma32(const uint8_t *luma32x32);
/* phash.rs:48-53 via image 0.25 Triangle resize + Rec.709 luma (recalled, UNVERIFIED) */
uint64_t orc_phash_image(const uint8_t *px, int layout, uint32_t w, uint32_t h, uint8_t *luma32_out);

/* ------------------------------------------------------------ Hamming ---- */

/* hamminghash.rs:55-58 / :34-36 */
uint32_t orc_hamming256(const uint8_t *a, const uint8_t *b);
uint32_t orc_hamming64(uint64_t a, uint64_t b);

/* hamminghash.rs:82-149.  width_bits is 64 or 256; hashes are n x (width/8) bytes
 * (u64 hashes in native little-endian). */
typedef struct orc_mih orc_mih;
orc_mih *orc_mih_new(const uint8_t *hashes, size_t n, int width_bits);
void orc_mih_free(orc_mih *);
/* bucket(chunk, value) -> pointer+len into the CSR values (hamminghash.rs:133-138) */
const uint32_t *orc_mih_bucket(const orc_mih *, int chunk, uint16_t value, size_t *len);
const uint32_t *orc_mih_offsets(const orc_mih *, size_t *len);

/* hamminghash.rs:191-271.  Output CSR: group_offsets has n_groups+1 entries.  Caller
 * frees with orc_free. */
int orc_find_groups(const orc_mih *, uint32_t max_dist, int threads, uint32_t **members,
                    uint32_t **group_offsets, size_t *n_groups);
void orc_free(void *);

/* scanner.rs:1640-1817 (edge phase through MIH + sequential union-find).
 *   hashes      n x 32 (rows without a hash are ignored)
 *   has_hash    n bytes or NULL (= all present)           scanner.rs:1658-1662
 *   variants    n x 8 x 32 or NULL; n_variants n bytes or NULL.  When variants is
 *               NULL every file queries with its own hash only (scanner.rs:1624-1627).
 *   low_conf    n bytes or NULL                            scanner.rs:1671,1699,1721
 *   similarity  must be <= 63 (returns -1 otherwise, mirroring the assert :1650-1655)
 * Outputs:
 *   out_label[i] = smallest index in i's connected component (canonical form of
 *                  the nondeterministically ordered groups_map)
 *   out_edge_count = edges.len() (scanner.rs:1778)
 *   edges_out / edges_cap: optional; first min(cap, count) edges in reference order
 *   use_mih: 1 = MIH probing exactly like the reference; 0 = brute-force all pairs
 */
int orc_group_generic(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                      const uint8_t *n_variants, const uint8_t *low_conf, size_t n,
                      uint32_t similarity, int threads, int use_mih, uint32_t *out_label,
                      uint64_t *out_edge_count, uint32_t *edges_out, size_t edges_cap);

/* bench.py timing aid: every chunk_stride-th 2000-file chunk of query files, its first sample_files files (0 = all) */
int orc_group_generic_sampled(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                              const uint8_t *n_variants, const uint8_t *low_conf, size_t n,
                              uint32_t similarity, int threads, size_t chunk_stride, size_t sample_files,
                              uint32_t *out_label, uint64_t *out_edge_count);

/* Same edge semantics restricted to the (row-block, col-block) tiles a rank owns;
 * used by the world_size-2 gloo tests to stand in for one GPU's tile kernel.
 * Writes a LOCAL parent forest (min-root) into out_parent. */
int orc_group_tiles_rank(const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                         const uint8_t *n_variants, const uint8_t *low_conf, size_t n,
                         uint32_t similarity, uint32_t tile, int rank, int world,
                         uint32_t *out_parent, uint64_t *out_edge_count);
/* merge `world` parent forests (world x n) into canonical labels */
void orc_merge_parents(const uint32_t *parents, int world, size_t n, uint32_t *out_label);

#ifdef __cplusplus
}
#endif
#endif
