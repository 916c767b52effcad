// hamming.cu -- hot path #2: all-pairs Hamming search under a threshold + device union-find.
//
// Replaces the edge phase and union-find of scanner::group_files_generic
// (scanner.rs:1640-1817) and HammingHash::hamming_distance (hamminghash.rs:34-36, :55-58).
// The reference finds the pairs through Multi-Index-Hashing bucket probes; for every allowed
// similarity (<= 63) the pigeonhole argument makes that set identical to the all-pairs set
// (SURVEY.md F3), which is what the tiles below enumerate.
//
// Kernel shape (hamming_tiles_kernel): persistent CTAs claim tiles of HT_TQ query rows x HT_TC
// candidates from a counter (in this GPU's memory, or -- multi-GPU group -- in one GPU's memory
// reached by every GPU with NVLink atomics: the GPUs steal work from one pool, so a slower GPU
// simply takes fewer tiles).  The candidates of a tile (8 x u32 each, 16 KB) are fetched by the
// bulk-copy engine (cp.async.bulk + mbarrier) into one of two shared-memory stages while the
// previous tile is being searched, and read with warp-broadcast LDS.128; every thread keeps HT_RQ
// query rows in registers, so one candidate fetch feeds HT_RQ pairs.
// A full distance costs 8 XOR + a 4-step carry-save compression (8 LOP3) + 4 POPC instead of 8 POPC:
// POPC issues at a quarter of the LOP3 rate, so trading POPCs for LOP3s balances the two pipes.
// The hot loop rarely needs it: the search is two-stage, and the first stage is a LOWER BOUND of the
// distance that costs one to three POPC (exact prefixes, or popc of an OR of several XOR words); only
// pairs whose bound is within the threshold get the full distance.  Which bound is used is decided on
// the device from a sampled selectivity (choose_variant); every choice gives identical results.
// Pairs under the threshold are rare; they take a divergent slow path that applies the exact
// edge rule (j > i, low-confidence => distance 0 only) and hooks the union-find.
//
// Nothing between the caller's buffers and the labels needs the host: the dense arrays, their
// sizes, the tile count and the claim counter are produced and consumed on the device
// (TileMeta), so a search is one stream of launches with a single synchronisation at the end.
#include <cub/device/device_scan.cuh>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "common.cuh"
#include "hamming_internal.cuh"
#include "tma.cuh"
#include "unionfind.cuh"

namespace {

constexpr int HT_THREADS = 256;
constexpr int HT_RQ = 4;                     // query rows per thread
constexpr int HT_TQ = HT_THREADS * HT_RQ;    // query rows per tile
constexpr int HT_TC = 512;                   // candidates per tile (16 KB per stage, two stages)
constexpr uint32_t NO_TILE = 0xFFFFFFFFu;

typedef unsigned long long u64;

// sizes known only on the device (files without a hash are dropped, variants per file vary)
struct TileMeta {
    uint32_t nc, nq;      // dense candidates / query rows
    uint32_t n_qb, n_cb;  // query blocks of HT_TQ rows, candidate blocks of HT_TC
    uint32_t pass96, pass128;   // of PF_SAMPLES sampled (query, candidate) pairs: prefix distance <= threshold
    uint32_t pass_or160, pass_or192, pass_or256;   // ... pairs whose OR lower bound over the first 160 / 192 / all 256 bits is <= threshold
    uint32_t pass_or64, pass_or128;                // ... whose ONE-POPC bound popc(x0 | x1) / popc(x0 | .. | x3) is <= threshold
};

// A first-stage bound only pays when it rejects almost every pair: a warp refines a candidate as soon as
// ONE of its 128 pairs survives, so at a survival rate of ~0.6 % a two-stage kernel already costs as
// much as the full-distance one (measured: bench.py worst_case).  The rate is a property of the input
// (uniform hashes, exact 96-bit prefix: 2.6e-4 at threshold 31; hashes sharing that prefix: 1), so it is
// sampled on the device for every bound and every CTA derives the same variant from the counters.
constexpr uint32_t PF_SAMPLES = 16384;
constexpr uint32_t PF_MAX_PASS = PF_SAMPLES * 3 / 1000;   // 0.3 %
constexpr uint32_t PF_MAX_PASS_1 = PF_SAMPLES / 2000;      // 0.05 % for the one-POPC bounds: a refinement costs more next to a cheaper hot loop

struct GroupArgs {
    const uint32_t *cand;     // [nc_pad][W] dense candidate hashes, zero padded to HT_TC rows
    const uint32_t *qry;      // [nq_pad][W] query rows ordered by file, 0xFF padded to HT_TQ rows
    const uint32_t *qfile;    // [nq_pad] dense file id of each query row (0xFFFFFFFF = padding)
    const uint8_t *lc;        // [nc] low-confidence flag per dense file, or nullptr
    uint32_t *parent;         // [nc] union-find forest over dense ids
    u64 *edge_count;
    uint2 *edges;             // optional edge sink (dense ids), capacity edges_cap
    u64 edges_cap;
    u64 *edges_n;
    const TileMeta *meta;
    const u64 *tile_start;    // [n_qb_max + 1] exclusive scan of the valid-tile count per query block
    uint32_t n_qb_max;        // tile_start[n_qb_max] = number of valid tiles
    u64 *next_tile;           // claim counter; claim c is tile c * claim_stride + claim_offset
    uint32_t claim_stride, claim_offset;
    int counter_is_remote;    // the counter may live in a peer GPU's memory (system-scope atomics)
    int force_pf;             // -1: from the sampled selectivity; 0 / 3 / 4: pinned (rh_ctx_set_option)
    uint32_t nc, threshold;   // nc is filled in from meta by the kernel
};

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// a | (b ^ c): one more word folded into an OR accumulator
__device__ __forceinline__ uint32_t or_xor(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// lower bound of the distance from the first 160 / 192 bits in two POPC: popc(x | y | z) <= popc(x) + popc(y) + popc(z),
// so  popc(x0 | x1 | x2) + popc(x3 | x4 [| x5]) <= d(first 160 / 192 bits) <= d.  For unrelated hashes an OR of
// three words has 28 +- 1.9 bits set, of two words 24 +- 2.4: the bound is ~N(52, 3.1^2) / ~N(56, 2.6^2), far more
// selective than the exact 96-bit prefix (48 +- 4.9) or the exact 128-bit prefix (64 +- 5.7, 3 POPC) at the same
// two POPC per pair; the 160-bit form needs 5 LOP3 like the 96-bit prefix, the 192-bit form 6.
// WORDS = 2 / 4: a single group, ONE POPC (24 +- 2.4 / 30 +- 1.4 bits for unrelated hashes): enough for strict
// thresholds (<= ~12 / <= ~24), where the search then runs at up to twice the two-POPC rate.
// WORDS = 7 stands for all 8 words in THREE groups (3 + 3 + 2 words, 3 POPC, ~N(80, 3.6^2)): selective up to the
// largest threshold the reference accepts (63), where the exact kernel needs 4 POPC and 16 LOP3 per pair.
template <int WORDS>
__device__ __forceinline__ uint32_t or_bound(const uint32_t (&q)[8], const uint4 &a, const uint4 &b) {
    if (WORDS == 2) return __popc(or_xor(q[0] ^ a.x, q[1], a.y));
    if (WORDS == 4) return __popc(or_xor(or_xor(or_xor(q[0] ^ a.x, q[1], a.y), q[2], a.z), q[3], a.w));
    const uint32_t o0 = or_xor(or_xor(q[0] ^ a.x, q[1], a.y), q[2], a.z);
    uint32_t o1 = or_xor(q[3] ^ a.w, q[4], b.x);
    if (WORDS >= 6) o1 = or_xor(o1, q[5], b.y);
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(__popc(o0)), "r"(__popc(o1)));   // the add on the FMA pipe
    if (WORDS == 7) asm("mad.lo.u32 %0, %1, 1, %0;" : "+r"(d) : "r"(__popc(or_xor(q[6] ^ b.z, q[7], b.w))));
    return d;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// popcount of a 256-bit XOR: carry-save adders fold 8 words into weights 1,1,2,4 -> 4 POPC.
//   p(x0..x2) = p(s0) + 2 p(c0);  p(x3..x5) = p(s1) + 2 p(c1);  p(s0,s1,x6) = p(s2) + 2 p(c2)
//   p(c0,c1,c2) = p(t0) + 2 p(d0)   =>  total = p(s2) + p(x7) + 2 p(t0) + 4 p(d0)
__device__ __forceinline__ uint32_t dist256(const uint32_t (&q)[8], const uint4 &a, const uint4 &b) {
    uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    uint32_t s0 = xor3(x0, x1, x2), c0 = maj3(x0, x1, x2);
    uint32_t s1 = xor3(x3, x4, x5), c1 = maj3(x3, x4, x5);
    uint32_t s2 = xor3(s0, s1, x6), c2 = maj3(s0, s1, x6);
    uint32_t t0 = xor3(c0, c1, c2), d0 = maj3(c0, c1, c2);
    return (__popc(s2) + __popc(x7)) + 2u * __popc(t0) + 4u * __popc(d0);
}

// Rare path, kept out of line so that none of its address arithmetic is hoisted into the hot
// loop.  `g` points at the CTA's shared-memory copy of the arguments.
__device__ __noinline__ uint32_t slow_hit(const GroupArgs *g, uint32_t d, uint32_t qf, uint32_t cj) {
    // exact edge rule of scanner.rs:1712-1724 in dense index space
    if (qf == 0xFFFFFFFFu || cj >= g->nc || cj <= qf) return 0;
    uint32_t lim = g->threshold;
    if (g->lc && (g->lc[qf] | g->lc[cj])) lim = 0;  // scanner.rs:1699, :1721
    if (d > lim) return 0;
    if (g->edges) {
        u64 k = atomicAdd(g->edges_n, 1ull);
        if (k < g->edges_cap) g->edges[k] = make_uint2(qf, cj);
    }
    RH_CHECK_IDX(qf, g->nc);
    rh::uf_unite(g->parent, qf, cj);
    return 1;
}

// first candidate block that can hold a pair with j > i for a query block whose smallest file id
// is f0 (rows are ordered by file; scanner.rs:1712-1714)
__device__ __forceinline__ uint32_t first_cand_block(uint32_t f0) { return (f0 + 1u) / (uint32_t)HT_TC; }

// ------------------------------------------------------------ tile scheduler ----
// Only tiles that can contain a pair with j > i exist: tile t (query block major, candidate block
// minor) is found from the exclusive scan `tile_start`.  Warp 0 of a CTA claims the next tile with
// one atomic, locates its query block with a 32-way search (every lane probes one point, a ballot
// picks the interval: 3 rounds instead of 15 dependent loads for 30 000 query blocks) and starts the
// bulk copy of its candidates.
struct TileSched {
    uint2 tile[2];                   // (query block, candidate block) staged in each buffer
    __align__(8) uint64_t bar[2];    // "candidates of stage s have landed"
};

__device__ __forceinline__ uint32_t find_query_block(const u64 *tile_start, uint32_t n_qb, u64 t, int lane) {
    uint32_t lo = 0, hi = n_qb;   // invariant: tile_start[lo] <= t, answer in [lo, hi)
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31u) >> 5;
        const uint32_t p = lo + (uint32_t)lane * step;
        const bool ok = p < hi && tile_start[p] <= t;   // monotone in the lane; lane 0 always holds
        const uint32_t k = 31u - (uint32_t)__clz((int)__ballot_sync(0xFFFFFFFFu, ok));
        lo += k * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

// warp 0, all lanes: claim a tile and describe it in sched.tile[buf]; returns its candidate block
// (NO_TILE when the pool is empty)
__device__ __forceinline__ uint2 claim_tile(const GroupArgs &g, const TileMeta &m, u64 n_tiles, int lane) {
    u64 c = 0;
    if (lane == 0) c = g.counter_is_remote ? atomicAdd_system(g.next_tile, 1ull) : atomicAdd(g.next_tile, 1ull);
    c = __shfl_sync(0xFFFFFFFFu, c, 0);
    const u64 t = c * g.claim_stride + g.claim_offset;
    if (t >= n_tiles) return make_uint2(NO_TILE, 0u);
    const uint32_t qb = find_query_block(g.tile_start, m.n_qb, t, lane);
    RH_CHECK_IDX(qb, m.n_qb);
    const uint32_t cb = first_cand_block(g.qfile[(size_t)qb * HT_TQ]) + (uint32_t)(t - g.tile_start[qb]);
    RH_CHECK_IDX(cb, m.n_cb);
    return make_uint2(qb, cb);
}

// PF = 0: every pair gets the full 256-bit distance (dist256).
// PF = 3 / 4: exact two-stage search.  The hot loop only measures the first PF words (96 / 128
// bits) -- a lower bound of the distance -- with one carry-save step (2 POPC for PF = 3); a pair
// whose partial distance already exceeds the threshold cannot be an edge.  For unrelated hashes
// the partial distance is ~N(16 PF, 8 PF), so at threshold <= 32 (PF = 3) / <= 46 (PF = 4) fewer
// than 1e-3 of the pairs survive; the survivors get the remaining words added and then take the
// same exact slow path.  Results are identical to PF = 0 for any input.
template <int PF>
__device__ __forceinline__ uint32_t search_tile(const uint4 *sc, int cn, uint32_t c0, const uint32_t (&q)[HT_RQ][8],
                                                const uint32_t (&qf)[HT_RQ], uint32_t T, const GroupArgs *s_g) {
    uint32_t local_edges = 0;
#pragma unroll (PF >= 1 && PF <= 3 ? 4 : 2)
    for (int c = 0; c < cn; c++) {
        const uint4 a = sc[2 * c];
        uint32_t d[HT_RQ];
        if (PF == 0) {
            const uint4 b = sc[2 * c + 1];
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) d[r] = dist256(q[r], a, b);
        } else if (PF == 1 || PF == 2) {   // one-POPC bounds over the first 128 / 64 bits (the second half is not read)
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) d[r] = or_bound<PF == 1 ? 4 : 2>(q[r], a, a);
        } else if (PF >= 5) {
            const uint4 b = sc[2 * c + 1];
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) d[r] = or_bound<PF>(q[r], a, b);
        } else {
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) {
                const uint32_t x0 = q[r][0] ^ a.x, x1 = q[r][1] ^ a.y, x2 = q[r][2] ^ a.z;
                // p(s) + 2 p(c) as one IMAD: keeps the add off the (busier) LOP3/IADD pipe
                asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(d[r]) : "r"(__popc(maj3(x0, x1, x2))), "r"(__popc(xor3(x0, x1, x2))));
                if (PF == 4) d[r] += __popc(q[r][3] ^ a.w);
            }
        }
        uint32_t mn = d[0];
#pragma unroll
        for (int r = 1; r < HT_RQ; r++) mn = min(mn, d[r]);
        if (mn <= T) {
            const uint4 b = sc[2 * c + 1];
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) {
                if (d[r] > T) continue;
                uint32_t full = d[r];
                if (PF >= 5 || PF == 1 || PF == 2) {
                    full = dist256(q[r], a, b);
                } else if (PF != 0) {
                    if (PF == 3) full += __popc(q[r][3] ^ a.w);
                    full += __popc(q[r][4] ^ b.x) + __popc(q[r][5] ^ b.y) + __popc(q[r][6] ^ b.z) +
                            __popc(q[r][7] ^ b.w);
                }
                if (full <= T) local_edges += slow_hit(s_g, full, qf[r], c0 + c);
            }
        }
    }
    return local_edges;
}

// the variant every CTA of every GPU derives from the sampled selectivity (or the pinned one); the host repeats
// the choice from the same counters for rh_hamming_last_variant.  Two-POPC bounds first, the cheaper first.
__host__ __device__ __forceinline__ int choose_variant(int force_pf, uint32_t threshold, const TileMeta &m) {
    if (force_pf >= 0) return force_pf;
    if (threshold > 63u) return 0;
    if (m.pass_or64 <= PF_MAX_PASS_1) return 2;
    if (m.pass_or128 <= PF_MAX_PASS_1) return 1;
    if (m.pass_or160 <= PF_MAX_PASS) return 5;
    if (m.pass96 <= PF_MAX_PASS) return 3;
    if (m.pass_or192 <= PF_MAX_PASS) return 6;
    if (m.pass_or256 <= PF_MAX_PASS) return 7;
    if (m.pass128 <= PF_MAX_PASS) return 4;
    return 0;
}
__device__ __forceinline__ int choose_prefilter(const GroupArgs &g, const TileMeta &m) {
    return choose_variant(g.force_pf, g.threshold, m);
}

// One instantiation per variant is launched for every search; the seven that the sampled selectivity
// did not choose return at once (~4 us each), so the variant is picked without a host round
// trip and each instantiation keeps its own lean register allocation.
template <int PF>
__global__ void __launch_bounds__(HT_THREADS, 4) hamming_tiles_kernel(const GroupArgs g) {
    __shared__ __align__(128) uint4 s_cand[2][HT_TC * 2];
    __shared__ GroupArgs s_g;
    __shared__ TileSched s_sched;
    const TileMeta m = *g.meta;
    if (choose_prefilter(g, m) != PF) return;
    const u64 n_tiles = g.tile_start[g.n_qb_max];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_g = g;
        s_g.nc = m.nc;
        rh::mbar_init(&s_sched.bar[0], 1);
        rh::mbar_init(&s_sched.bar[1], 1);
        rh::mbar_fence_init();
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = HT_TC * 32;
    const uint4 *cand4 = reinterpret_cast<const uint4 *>(g.cand);
    auto stage = [&](int buf) {   // warp 0
        const uint2 t = claim_tile(g, m, n_tiles, lane);
        if (lane == 0) {
            s_sched.tile[buf] = t;
            if (t.x != NO_TILE) {
                rh::mbar_arrive_expect_tx(&s_sched.bar[buf], STAGE_BYTES);
                rh::bulk_g2s(s_cand[buf], cand4 + (size_t)t.y * (HT_TC * 2), STAGE_BYTES, &s_sched.bar[buf]);
            }
        }
    };
    if (warp == 0) stage(0);
    __syncthreads();

    uint32_t q[HT_RQ][8];
    uint32_t qf[HT_RQ];
    uint32_t cur_qb = NO_TILE;
    const uint32_t T = g.threshold;
    uint32_t local_edges = 0;
    for (uint32_t it = 0;; it++) {
        const int buf = (int)(it & 1u);
        const uint2 tile = s_sched.tile[buf];
        if (tile.x == NO_TILE) break;
        // the other stage was last read in the previous iteration, which ended with a barrier
        if (warp == 0) stage(buf ^ 1);
        if (tile.x != cur_qb) {
            cur_qb = tile.x;
#pragma unroll
            for (int r = 0; r < HT_RQ; r++) {
                const uint32_t row = cur_qb * HT_TQ + r * HT_THREADS + threadIdx.x;
                RH_CHECK_IDX(row, (size_t)m.n_qb * HT_TQ);
                const uint4 *p = reinterpret_cast<const uint4 *>(g.qry) + (size_t)row * 2;
                uint4 a = p[0], b = p[1];
                q[r][0] = a.x; q[r][1] = a.y; q[r][2] = a.z; q[r][3] = a.w;
                q[r][4] = b.x; q[r][5] = b.y; q[r][6] = b.z; q[r][7] = b.w;
                qf[r] = g.qfile[row];
            }
        }
        const uint32_t c0 = tile.y * HT_TC;
        const int cn = (int)(min(c0 + (uint32_t)HT_TC, m.nc) - c0);
        rh::mbar_wait(&s_sched.bar[buf], (it >> 1) & 1u);
        const uint4 *sc = s_cand[buf];
        local_edges += search_tile<PF>(sc, cn, c0, q, qf, T, &s_g);
        __syncthreads();   // everyone is done with this stage; the next tile's description is visible
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_edges += __shfl_xor_sync(0xFFFFFFFFu, local_edges, o);
    if (lane == 0 && local_edges) atomicAdd(g.edge_count, (u64)local_edges);
}

// 64-bit hashes (hamminghash.rs:23-41): 2 words per hash, the same persistent tile pool; the 4 KB of
// candidates per tile are loaded directly.  Throughput is not a target here (no reference caller
// groups u64 hashes).
__global__ void __launch_bounds__(HT_THREADS) hamming_tiles_u64_kernel(const GroupArgs g) {
    __shared__ uint2 s_cand[HT_TC];
    __shared__ GroupArgs s_g;
    __shared__ uint2 s_tile;
    const TileMeta m = *g.meta;
    const u64 n_tiles = g.tile_start[g.n_qb_max];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_g = g;
        s_g.nc = m.nc;
    }
    const uint32_t T = g.threshold;
    const bool use_bound = g.force_pf != 0 && m.pass_or160 <= PF_MAX_PASS;   // sampled like the 256-bit variants
    uint32_t local_edges = 0;
    for (;;) {
        __syncthreads();   // the previous tile's candidates and description are no longer needed
        if (warp == 0) {
            const uint2 t = claim_tile(g, m, n_tiles, lane);
            if (lane == 0) s_tile = t;
        }
        __syncthreads();
        const uint2 tile = s_tile;
        if (tile.x == NO_TILE) break;
        const uint32_t q0 = tile.x * HT_TQ, c0 = tile.y * HT_TC;
        const int cn = (int)(min(c0 + (uint32_t)HT_TC, m.nc) - c0);
        const uint2 *cand2 = reinterpret_cast<const uint2 *>(g.cand) + c0;
        for (int i = threadIdx.x; i < HT_TC; i += HT_THREADS) s_cand[i] = cand2[i];
        uint2 q[HT_RQ];
        uint32_t qf[HT_RQ];
#pragma unroll
        for (int r = 0; r < HT_RQ; r++) {
            const uint32_t row = q0 + r * HT_THREADS + threadIdx.x;
            q[r] = reinterpret_cast<const uint2 *>(g.qry)[row];
            qf[r] = g.qfile[row];
        }
        __syncthreads();
        if (use_bound) {
            // one POPC per pair: popc(x | y) <= popc(x) + popc(y) (24 +- 2.4 bits for unrelated values, thresholds
            // are <= 15); the pairs it lets through get the exact distance
            for (int c = 0; c < cn; c++) {
                const uint2 a = s_cand[c];
                uint32_t ob[HT_RQ];
#pragma unroll
                for (int r = 0; r < HT_RQ; r++) ob[r] = __popc(or_xor(q[r].x ^ a.x, q[r].y, a.y));
                uint32_t mn = ob[0];
#pragma unroll
                for (int r = 1; r < HT_RQ; r++) mn = min(mn, ob[r]);
                if (mn <= T) {
#pragma unroll
                    for (int r = 0; r < HT_RQ; r++) {
                        if (ob[r] > T) continue;
                        const uint32_t d = __popc(q[r].x ^ a.x) + __popc(q[r].y ^ a.y);
                        if (d <= T) local_edges += slow_hit(&s_g, d, qf[r], c0 + c);
                    }
                }
            }
        } else {
            for (int c = 0; c < cn; c++) {
                const uint2 a = s_cand[c];
#pragma unroll
                for (int r = 0; r < HT_RQ; r++) {
                    uint32_t d = __popc(q[r].x ^ a.x) + __popc(q[r].y ^ a.y);
                    if (d <= T) local_edges += slow_hit(&s_g, d, qf[r], c0 + c);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_edges += __shfl_xor_sync(0xFFFFFFFFu, local_edges, o);
    if (lane == 0 && local_edges) atomicAdd(g.edge_count, (u64)local_edges);
}

// ------------------------------------------------------------ preparation ----

// valid[i] (file has a hash) and nv[i] (number of query rows of file i); entry n is 0 so that
// the exclusive scans end with the totals.
__global__ void flags_kernel(const uint8_t *has_hash, const uint8_t *n_variants, int has_variants,
                             uint32_t n, uint32_t *valid, uint32_t *nv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t v = 0, q = 0;
    if (i < n) {
        v = has_hash ? (has_hash[i] != 0) : 1u;
        if (v) q = has_variants ? (n_variants ? min(max((uint32_t)n_variants[i], 1u), 8u) : 8u) : 1u;
    }
    valid[i] = v;
    nv[i] = q;
}

// the scans' totals -> TileMeta; zeroes the edge counter, the edge-list fill and the claim counter
__global__ void meta_kernel(const uint32_t *dpos, const uint32_t *qoff, uint32_t n, TileMeta *meta, u64 *counters) {
    const uint32_t nc = dpos[n], nq = qoff[n];
    meta->nc = nc;
    meta->nq = nq;
    meta->n_qb = (nq + HT_TQ - 1) / HT_TQ;
    meta->n_cb = (nc + HT_TC - 1) / HT_TC;
    meta->pass96 = 0;
    meta->pass128 = 0;
    meta->pass_or160 = 0;
    meta->pass_or192 = 0;
    meta->pass_or256 = 0;
    meta->pass_or64 = 0;
    meta->pass_or128 = 0;
    counters[0] = 0;   // comparison_count
    counters[1] = 0;   // edges written to the optional sink
    counters[2] = 0;   // next tile to claim (unused when the group's shared counter is given)
}

__device__ __forceinline__ uint32_t load_le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// One thread per (file, word): scatters candidate words, query rows, ids and flags into the
// dense arrays.  Bytes are read one by one, so the caller's buffers need no alignment.
template <int W>
__global__ void scatter_kernel(const uint8_t *hashes, const uint8_t *variants, const uint8_t *low_conf,
                               uint32_t n, const uint32_t *valid, const uint32_t *dpos,
                               const uint32_t *nv, const uint32_t *qoff, uint32_t *cand,
                               uint32_t *cand_idx, uint8_t *lc, uint32_t *qry, uint32_t *qfile) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const uint32_t i = (uint32_t)(t / W), w = (uint32_t)(t % W);
    if (i >= n || !valid[i]) return;
    const uint32_t k = dpos[i];
    cand[(size_t)k * W + w] = load_le32(hashes + (size_t)i * (W * 4) + w * 4);
    if (w == 0) {
        cand_idx[k] = i;
        if (lc) lc[k] = low_conf[i] ? 1 : 0;
    }
    const uint32_t cnt = nv[i], qo = qoff[i];
    for (uint32_t v = 0; v < cnt; v++) {
        const uint8_t *src = variants ? variants + ((size_t)i * 8 + v) * (W * 4) : hashes + (size_t)i * (W * 4);
        qry[(size_t)(qo + v) * W + w] = load_le32(src + w * 4);
        if (w == 0) qfile[qo + v] = k;
    }
}

// the rows between the dense totals and the next block boundary: candidates zero, query rows 0xFF
// (their qfile = 0xFFFFFFFF makes slow_hit reject them), so that tiles never need a row mask
template <int W>
__global__ void pad_kernel(const TileMeta *meta, uint32_t *cand, uint32_t *qry, uint32_t *qfile) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = t / W, w = t % W;
    const TileMeta m = *meta;
    if (row < (uint32_t)HT_TQ) {
        const size_t r = (size_t)m.nq + row;
        if (r < (size_t)m.n_qb * HT_TQ) {
            qry[r * W + w] = 0xFFFFFFFFu;
            if (w == 0) qfile[r] = 0xFFFFFFFFu;
        }
    }
    if (row < (uint32_t)HT_TC) {
        const size_t r = (size_t)m.nc + row;
        if (r < (size_t)m.n_cb * HT_TC) cand[r * W + w] = 0u;
    }
}

// prefix-filter selectivity on PF_SAMPLES pseudo-random (query row, candidate) pairs
// (a query row is never sampled against its own file: in a small input those pairs alone would exceed the limit)
__global__ void selectivity_kernel(const uint32_t *cand, const uint32_t *qry, const uint32_t *qfile, TileMeta *meta,
                                   uint32_t threshold) {
    const TileMeta m = *meta;
    if (m.nc == 0 || m.nq == 0) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = t * 2654435761u + 0x9E3779B9u;
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    uint32_t y = x * 0x27D4EB2Fu + 0x165667B1u;
    y ^= y >> 15; y *= 0x2C1B3C6Du; y ^= y >> 12;
    const uint32_t qi = x % m.nq;
    uint32_t ci = y % m.nc;
    if (ci == qfile[qi]) ci = (ci + 1u) % m.nc;
    const uint4 a = reinterpret_cast<const uint4 *>(qry)[(size_t)qi * 2];
    const uint4 b = reinterpret_cast<const uint4 *>(cand)[(size_t)ci * 2];
    const uint32_t d96 = __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z);
    const uint32_t d128 = d96 + __popc(a.w ^ b.w);
    const uint4 a2 = reinterpret_cast<const uint4 *>(qry)[(size_t)qi * 2 + 1];
    const uint4 b2 = reinterpret_cast<const uint4 *>(cand)[(size_t)ci * 2 + 1];
    const uint32_t o0 = __popc((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z));
    const uint32_t dor5 = o0 + __popc((a.w ^ b.w) | (a2.x ^ b2.x));
    const uint32_t dor = o0 + __popc((a.w ^ b.w) | (a2.x ^ b2.x) | (a2.y ^ b2.y));
    const uint32_t p96 = __syncthreads_count(d96 <= threshold), p128 = __syncthreads_count(d128 <= threshold);
    const uint32_t dor7 = dor + __popc((a2.z ^ b2.z) | (a2.w ^ b2.w));
    const uint32_t por = __syncthreads_count(dor <= threshold), por5 = __syncthreads_count(dor5 <= threshold);
    const uint32_t por7 = __syncthreads_count(dor7 <= threshold);
    const uint32_t por2 = __syncthreads_count(__popc((a.x ^ b.x) | (a.y ^ b.y)) <= threshold);
    const uint32_t por4 = __syncthreads_count(__popc((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w)) <= threshold);
    if (threadIdx.x == 0) {
        if (p96) atomicAdd(&meta->pass96, p96);
        if (p128) atomicAdd(&meta->pass128, p128);
        if (por) atomicAdd(&meta->pass_or192, por);
        if (por5) atomicAdd(&meta->pass_or160, por5);
        if (por7) atomicAdd(&meta->pass_or256, por7);
        if (por2) atomicAdd(&meta->pass_or64, por2);
        if (por4) atomicAdd(&meta->pass_or128, por4);
    }
}

__global__ void selectivity_u64_kernel(const uint32_t *cand, const uint32_t *qry, const uint32_t *qfile, TileMeta *meta,
                                       uint32_t threshold) {
    const TileMeta m = *meta;
    if (m.nc == 0 || m.nq == 0) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = t * 2654435761u + 0x9E3779B9u;
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    uint32_t y = x * 0x27D4EB2Fu + 0x165667B1u;
    y ^= y >> 15; y *= 0x2C1B3C6Du; y ^= y >> 12;
    const uint32_t qi = x % m.nq;
    uint32_t ci = y % m.nc;
    if (ci == qfile[qi]) ci = (ci + 1u) % m.nc;
    const uint2 a = reinterpret_cast<const uint2 *>(qry)[qi];
    const uint2 b = reinterpret_cast<const uint2 *>(cand)[ci];
    const uint32_t p = __syncthreads_count(__popc((a.x ^ b.x) | (a.y ^ b.y)) <= threshold);
    if (threadIdx.x == 0 && p) atomicAdd(&meta->pass_or160, p);
}

// number of candidate blocks that can hold a pair with j > i, per query block (entries from
// meta.n_qb to n_qb_max are 0, so the exclusive scan ends with the total)
__global__ void tile_counts_kernel(const uint32_t *qfile, const TileMeta *meta, uint32_t n_qb_max, u64 *counts) {
    const uint32_t qb = blockIdx.x * blockDim.x + threadIdx.x;
    if (qb > n_qb_max) return;
    const TileMeta m = *meta;
    u64 c = 0;
    if (qb < m.n_qb) {
        const uint32_t f0 = qfile[(size_t)qb * HT_TQ];
        if (f0 != 0xFFFFFFFFu && f0 + 1u < m.nc) c = m.n_cb - min(m.n_cb, first_cand_block(f0));
    }
    counts[qb] = c;
}

__global__ void iota_kernel(uint32_t *p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// label of sparse file i = sparse index of the root of its dense id (files without a hash
// are their own label, scanner.rs:1658-1662 keeps them out of the search).
__global__ void labels_kernel(uint32_t *parent, const uint32_t *valid, const uint32_t *dpos,
                              const uint32_t *cand_idx, uint32_t n, uint32_t *out_label) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_label[i] = valid[i] ? cand_idx[rh::uf_find(parent, dpos[i])] : i;
}

__global__ void merge_kernel(const uint32_t *parents, uint32_t n, int world, uint32_t *parent) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)n * world) return;
    uint32_t i = (uint32_t)(t % n);
    uint32_t p = parents[t];
    if (p != i && p < n) rh::uf_unite(parent, i, p);
}

__global__ void flatten_kernel(uint32_t *parent, uint32_t n, uint32_t *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rh::uf_find(parent, i);
}

template <int W>
__global__ void distances_kernel(const uint8_t *a, const uint8_t *b, size_t n, uint32_t *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t d = 0;
#pragma unroll
    for (int w = 0; w < W; w++)
        d += __popc(load_le32(a + i * (W * 4) + w * 4) ^ load_le32(b + i * (W * 4) + w * 4));
    out[i] = d;
}

__global__ void edges_to_sparse_kernel(uint2 *edges, u64 cap, const u64 *n_edges, const uint32_t *cand_idx) {
    u64 k = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 m = *n_edges < cap ? *n_edges : cap;
    if (k >= m) return;
    uint2 e = edges[k];
    edges[k] = make_uint2(cand_idx[e.x], cand_idx[e.y]);
}

// max over a group's members of (min over the pivot's variants of the Hamming distance):
// analyze_group_with_features' max_dist (scanner.rs:2217-2241).  One thread per member.
__global__ void group_max_dist_kernel(const uint8_t *pivots, const uint8_t *n_pivot_variants, const uint8_t *members,
                                      const uint32_t *member_group, size_t m, uint32_t *out_max) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t g = member_group[i];
    uint32_t h[8];
#pragma unroll
    for (int w = 0; w < 8; w++) h[w] = load_le32(members + i * 32 + w * 4);
    const int nv = n_pivot_variants ? max(1, min((int)n_pivot_variants[g], 8)) : 8;
    uint32_t best = 0xFFFFFFFFu;
    for (int v = 0; v < nv; v++) {
        const uint8_t *pv = pivots + ((size_t)g * 8 + v) * 32;
        uint32_t d = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(h[w] ^ load_le32(pv + w * 4));
        best = min(best, d);
    }
    atomicMax(out_max + g, best);
}

struct Prepared {
    GroupArgs g;
    const uint32_t *valid, *dpos, *cand_idx;
    u64 *counters;
    size_t tiles_upper;   // host-side upper bound of the tile count (sizes the persistent grid)
};

// Builds the dense candidate / query arrays on the device; every size the host needs is an upper
// bound derived from n (the exact totals stay on the device in TileMeta).  W = words per hash.
template <int W>
int prepare(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
            const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
            const rh::HammingPlan &plan, Prepared *out) {
    using namespace rh;
    cudaStream_t st = ctx->stream;
    const size_t hb = W * 4;
    const uint8_t *d_hashes, *d_has, *d_var, *d_nv, *d_lc;
    RH_TRY(stage_in(ctx, hashes, (size_t)n * hb, S_IN0, &d_hashes));
    RH_TRY(stage_in(ctx, has_hash, (size_t)n, S_IN1, &d_has));
    RH_TRY(stage_in(ctx, variants, (size_t)n * 8 * hb, S_IN2, &d_var));
    RH_TRY(stage_in(ctx, n_variants, (size_t)n, S_IN3, &d_nv));
    RH_TRY(stage_in(ctx, low_conf, (size_t)n, S_IN4, &d_lc));

    const size_t nq_max = (size_t)n * (d_var ? 8 : 1);
    const size_t nc_pad = (size_t)cdiv((size_t)n, HT_TC) * HT_TC;
    const size_t nq_pad = (size_t)cdiv(nq_max, HT_TQ) * HT_TQ;
    const uint32_t n_qb_max = (uint32_t)(nq_pad / HT_TQ), n_cb_max = (uint32_t)(nc_pad / HT_TC);

    uint32_t *valid, *nv, *dpos, *qoff;
    void *p;
    RH_TRY(scratch(ctx, S_W0, (size_t)(n + 1) * 4, &p)); valid = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W1, (size_t)(n + 1) * 4, &p)); nv = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W2, (size_t)(n + 1) * 4, &p)); dpos = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W3, (size_t)(n + 1) * 4, &p)); qoff = (uint32_t *)p;
    uint32_t *cand, *cand_idx, *qry, *qfile, *parent;
    uint8_t *lc = nullptr;
    u64 *counters, *tcount, *tstart;
    RH_TRY(scratch(ctx, S_W5, nc_pad * hb, &p)); cand = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W6, nc_pad * 4, &p)); cand_idx = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W7, nq_pad * hb, &p)); qry = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W8, nq_pad * 4, &p)); qfile = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W9, nc_pad * 4, &p)); parent = (uint32_t *)p;
    RH_TRY(scratch(ctx, S_W10, 128, &p)); counters = (u64 *)p;
    TileMeta *meta = reinterpret_cast<TileMeta *>(counters + 8);
    if (d_lc) {
        RH_TRY(scratch(ctx, S_W11, nc_pad, &p));
        lc = (uint8_t *)p;
    }
    RH_TRY(scratch(ctx, S_W12, (size_t)(n_qb_max + 1) * 8, &p)); tcount = (u64 *)p;
    RH_TRY(scratch(ctx, S_W13, (size_t)(n_qb_max + 1) * 8, &p)); tstart = (u64 *)p;
    size_t tmp_a = 0, tmp_b = 0;
    RH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, valid, dpos, (int)(n + 1), st));
    RH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, tcount, tstart, (int)(n_qb_max + 1), st));
    const size_t tmp_bytes = tmp_a > tmp_b ? tmp_a : tmp_b;
    void *tmp;
    RH_TRY(scratch(ctx, S_W4, tmp_bytes, &tmp));

    flags_kernel<<<cdiv(n + 1, 256), 256, 0, st>>>(d_has, d_nv, d_var != nullptr, (uint32_t)n, valid, nv);
    RH_LAUNCHED(ctx, "flags_kernel");
    size_t tb = tmp_bytes;
    RH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, valid, dpos, (int)(n + 1), st));
    tb = tmp_bytes;
    RH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, nv, qoff, (int)(n + 1), st));
    ctx->launches += 2;
    meta_kernel<<<1, 1, 0, st>>>(dpos, qoff, (uint32_t)n, meta, counters);
    RH_LAUNCHED(ctx, "meta_kernel");
    scatter_kernel<W><<<cdiv((size_t)n * W, 256), 256, 0, st>>>(d_hashes, d_var, d_lc, (uint32_t)n, valid, dpos, nv, qoff,
                                                                cand, cand_idx, lc, qry, qfile);
    RH_LAUNCHED(ctx, "scatter_kernel");
    pad_kernel<W><<<cdiv((size_t)HT_TQ * W, 256), 256, 0, st>>>(meta, cand, qry, qfile);
    RH_LAUNCHED(ctx, "pad_kernel");
    iota_kernel<<<cdiv(nc_pad, 256), 256, 0, st>>>(parent, (uint32_t)nc_pad);
    RH_LAUNCHED(ctx, "iota_kernel");
    if (W == 8) {
        selectivity_kernel<<<PF_SAMPLES / 256, 256, 0, st>>>(cand, qry, qfile, meta, similarity);
        RH_LAUNCHED(ctx, "selectivity_kernel");
    } else {
        selectivity_u64_kernel<<<PF_SAMPLES / 256, 256, 0, st>>>(cand, qry, qfile, meta, similarity);
        RH_LAUNCHED(ctx, "selectivity_u64_kernel");
    }
    tile_counts_kernel<<<cdiv(n_qb_max + 1, 256), 256, 0, st>>>(qfile, meta, n_qb_max, tcount);
    RH_LAUNCHED(ctx, "tile_counts_kernel");
    tb = tmp_bytes;
    RH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, tcount, tstart, (int)(n_qb_max + 1), st));
    ctx->launches += 1;

    GroupArgs &g = out->g;
    g.cand = cand;
    g.qry = qry;
    g.qfile = qfile;
    g.lc = lc;
    g.parent = parent;
    g.edge_count = counters;
    g.edges = nullptr;
    g.edges_cap = 0;
    g.edges_n = counters + 1;
    g.meta = meta;
    g.tile_start = tstart;
    g.n_qb_max = n_qb_max;
    if (plan.shared_counter) {   // one pool for every GPU of the group
        g.next_tile = plan.shared_counter;
        g.claim_stride = 1;
        g.claim_offset = 0;
        g.counter_is_remote = 1;
    } else {                     // static cyclic ownership: tile t belongs to rank t mod world
        g.next_tile = counters + 2;
        g.claim_stride = (uint32_t)plan.world;
        g.claim_offset = (uint32_t)plan.rank;
        g.counter_is_remote = 0;
    }
    g.force_pf = ctx->force_prefilter;
    g.nc = 0;
    g.threshold = similarity;
    out->valid = valid;
    out->dpos = dpos;
    out->cand_idx = cand_idx;
    out->counters = counters;
    out->tiles_upper = (size_t)n_qb_max * n_cb_max;
    return RH_OK;
}

template <int W>
int run_tiles(rh_ctx *ctx, const Prepared &pr, int world) {
    cudaStream_t st = ctx->stream;
    ctx->last_ms = 0.0;
    ctx->last_units = 0.0;
    // persistent grid: 4 CTAs per SM, fewer when the whole pool is smaller than that
    size_t share = pr.tiles_upper / (size_t)(world > 0 ? world : 1) + 1;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4, share);
    RH_CUDA(ctx, cudaEventRecord(ctx->ev_a, st));
    if (W == 8) {
        // the variant is chosen on the device (choose_prefilter): the other two launches return at once
        const int f = pr.g.force_pf;
        if (f < 0 || f == 3) hamming_tiles_kernel<3><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 2) hamming_tiles_kernel<2><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 1) hamming_tiles_kernel<1><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 5) hamming_tiles_kernel<5><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 6) hamming_tiles_kernel<6><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 7) hamming_tiles_kernel<7><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 4) hamming_tiles_kernel<4><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0 || f == 0) hamming_tiles_kernel<0><<<grid, HT_THREADS, 0, st>>>(pr.g);
        if (f < 0) ctx->launches += 7;   // eight instantiations launched, RH_LAUNCHED below counts one
    } else
        hamming_tiles_u64_kernel<<<grid, HT_THREADS, 0, st>>>(pr.g);
    RH_LAUNCHED(ctx, "hamming_tiles_kernel");
    RH_CUDA(ctx, cudaEventRecord(ctx->ev_b, st));
    ctx->timing_pending = true;
    return RH_OK;
}

int check_group_args(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *variants, const uint8_t *n_variants, int64_t n,
                     uint32_t similarity, uint32_t max_similarity, int rank, int world) {
    using namespace rh;
    if (n < 0 || n > 0x7FFFFFF0ll) return fail(ctx, RH_EINVAL, "n out of range");
    if (similarity > max_similarity) return fail(ctx, RH_EINVAL, "similarity above 63 (scanner.rs:1650-1655)");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, RH_EINVAL, "bad rank/world");
    if ((n > 0 && !hashes) || (n_variants && !variants)) return fail(ctx, RH_EINVAL, "null hashes");
    // dense ids, query rows and their scans are 32-bit
    if ((uint64_t)n * (variants ? 8u : 1u) > 0xFFFFF000ull)
        return fail(ctx, RH_EUNSUPPORTED, "more than 2^32 query rows (n x variants): split the library");
    return RH_OK;
}

template <int W>
int enqueue_impl(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                 const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                 const rh::HammingPlan &plan, uint32_t *d_out_label, uint2 *d_edges, size_t edges_cap,
                 const u64 **d_counters) {
    using namespace rh;
    cudaStream_t st = ctx->stream;
    Prepared pr;
    RH_TRY(prepare<W>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, plan, &pr));
    if (d_edges && edges_cap) {
        pr.g.edges = d_edges;
        pr.g.edges_cap = edges_cap;
    }
    if (plan.before_tiles) RH_CUDA(ctx, cudaStreamWaitEvent(st, plan.before_tiles, 0));
    RH_TRY(run_tiles<W>(ctx, pr, plan.shared_counter ? plan.world : plan.world));
    if (d_out_label) {
        labels_kernel<<<cdiv(n, 256), 256, 0, st>>>(pr.g.parent, pr.valid, pr.dpos, pr.cand_idx, (uint32_t)n, d_out_label);
        RH_LAUNCHED(ctx, "labels_kernel");
    }
    if (d_edges && edges_cap) {
        // dense ids -> file indices (the edge list itself is unordered)
        edges_to_sparse_kernel<<<cdiv(edges_cap, 256), 256, 0, st>>>(pr.g.edges, pr.g.edges_cap, pr.g.edges_n, pr.cand_idx);
        RH_LAUNCHED(ctx, "edges_to_sparse_kernel");
    }
    if (d_counters) *d_counters = pr.counters;
    return RH_OK;
}

template <int W>
int group_impl(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
               const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
               uint32_t max_similarity, int rank, int world, uint32_t *out_label, uint64_t *out_edge_count,
               uint32_t *out_edges, size_t edges_cap) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    RH_TRY(check_group_args(ctx, hashes, variants, n_variants, n, similarity, max_similarity, rank, world));
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (out_edge_count) *out_edge_count = 0;
    if (n == 0) return RH_OK;
    OutBuf<uint32_t> lab, edg;
    RH_TRY(lab.prepare(ctx, out_label, (size_t)n, S_OUT0));
    if (out_edges && edges_cap) RH_TRY(edg.prepare(ctx, out_edges, edges_cap * 2, S_OUT1));
    HammingPlan plan;
    plan.rank = rank;
    plan.world = world;
    const u64 *d_counters = nullptr;
    RH_TRY(enqueue_impl<W>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, plan, lab.dev,
                           reinterpret_cast<uint2 *>(edg.dev), edg.dev ? edges_cap : 0, &d_counters));
    RH_TRY(lab.finish(ctx));
    RH_TRY(edg.finish(ctx));
    u64 counts[16] = {0};   // [0] comparison count, [1] edges written, [8..] the TileMeta the kernels used
    static_assert(sizeof(TileMeta) <= 64, "counters + meta share one 128-byte scratch block");
    RH_CUDA(ctx, cudaMemcpyAsync(counts, d_counters, 128, cudaMemcpyDeviceToHost, st));
    RH_CUDA(ctx, cudaStreamSynchronize(st));   // the only host synchronisation of a search
    RH_TRY(finish_timing(ctx));
    if (out_edge_count) *out_edge_count = counts[0];
    TileMeta hm;
    memcpy(&hm, counts + 8, sizeof(hm));
    // (u64 hashes: 1 = the one-POPC OR bound, 0 = exact distance for every pair)
    ctx->last_hamming_variant = W == 8 ? choose_variant(ctx->force_prefilter, similarity, hm)
                                       : (ctx->force_prefilter != 0 && hm.pass_or160 <= PF_MAX_PASS ? 1 : 0);
    return RH_OK;
}

}  // namespace

namespace rh {

int finish_timing(rh_ctx *ctx) {
    if (ctx->timing_pending) {
        float ms = 0.f;
        RH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        ctx->last_ms = ms;
        ctx->timing_pending = false;
    }
    return RH_OK;
}

int hamming_group_enqueue(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                          const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                          const HammingPlan &plan, uint32_t *d_out_label, const unsigned long long **d_counters) {
    RH_TRY(check_group_args(ctx, hashes, variants, n_variants, n, similarity, RH_MAX_SIMILARITY_256, plan.rank, plan.world));
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    return enqueue_impl<8>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, plan, d_out_label, nullptr,
                           0, d_counters);
}

int uf_merge_enqueue(rh_ctx *ctx, const uint32_t *d_parents, int world, int64_t n, uint32_t *d_out_label) {
    cudaStream_t st = ctx->stream;
    void *p;
    RH_TRY(scratch(ctx, S_W9, (size_t)n * 4, &p));
    uint32_t *parent = (uint32_t *)p;
    iota_kernel<<<cdiv(n, 256), 256, 0, st>>>(parent, (uint32_t)n);
    RH_LAUNCHED(ctx, "iota_kernel");
    merge_kernel<<<cdiv((size_t)n * world, 256), 256, 0, st>>>(d_parents, (uint32_t)n, world, parent);
    RH_LAUNCHED(ctx, "merge_kernel");
    flatten_kernel<<<cdiv(n, 256), 256, 0, st>>>(parent, (uint32_t)n, d_out_label);
    RH_LAUNCHED(ctx, "flatten_kernel");
    return RH_OK;
}

}  // namespace rh

extern "C" {

int rh_hamming_group(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                     const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                     uint32_t *out_label, uint64_t *out_edge_count) {
    return group_impl<8>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, RH_MAX_SIMILARITY_256, 0, 1,
                         out_label, out_edge_count, nullptr, 0);
}

int rh_hamming_group_shard(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                           const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                           int rank, int world, uint32_t *out_parent, uint64_t *out_edge_count) {
    return group_impl<8>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, RH_MAX_SIMILARITY_256, rank,
                         world, out_parent, out_edge_count, nullptr, 0);
}

int rh_hamming_group_u64(rh_ctx *ctx, const uint64_t *hashes, const uint8_t *has_hash, const uint64_t *variants,
                         const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                         uint32_t *out_label, uint64_t *out_edge_count) {
    return group_impl<2>(ctx, reinterpret_cast<const uint8_t *>(hashes), has_hash,
                         reinterpret_cast<const uint8_t *>(variants), n_variants, low_conf, n, similarity,
                         RH_MAX_SIMILARITY_256, 0, 1, out_label, out_edge_count, nullptr, 0);
}

int rh_uf_merge(rh_ctx *ctx, const uint32_t *parents, int world, int64_t n, uint32_t *out_label) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    if (n < 0 || n > 0x7FFFFFF0ll || world < 1 || (n > 0 && (!parents || !out_label)))
        return fail(ctx, RH_EINVAL, "rh_uf_merge: bad arguments");
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t *d_par;
    RH_TRY(stage_in(ctx, parents, (size_t)n * world, S_IN0, &d_par));
    OutBuf<uint32_t> lab;
    RH_TRY(lab.prepare(ctx, out_label, (size_t)n, S_OUT0));
    RH_TRY(uf_merge_enqueue(ctx, d_par, world, n, lab.dev));
    RH_TRY(lab.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RH_OK;
}

int rh_hamming_distances(rh_ctx *ctx, const uint8_t *a, const uint8_t *b, int64_t n, uint32_t *out) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(ctx, RH_EINVAL, "rh_hamming_distances: bad arguments");
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t *da, *db;
    RH_TRY(stage_in(ctx, a, (size_t)n * 32, S_IN0, &da));
    RH_TRY(stage_in(ctx, b, (size_t)n * 32, S_IN1, &db));
    OutBuf<uint32_t> o;
    RH_TRY(o.prepare(ctx, out, (size_t)n, S_OUT0));
    distances_kernel<8><<<cdiv(n, 256), 256, 0, st>>>(da, db, (size_t)n, o.dev);
    RH_LAUNCHED(ctx, "distances_kernel");
    RH_TRY(o.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}

int rh_hamming_distances_u64(rh_ctx *ctx, const uint64_t *a, const uint64_t *b, int64_t n, uint32_t *out) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(ctx, RH_EINVAL, "rh_hamming_distances_u64: bad arguments");
    if (n == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t *da, *db;
    RH_TRY(stage_in(ctx, reinterpret_cast<const uint8_t *>(a), (size_t)n * 8, S_IN0, &da));
    RH_TRY(stage_in(ctx, reinterpret_cast<const uint8_t *>(b), (size_t)n * 8, S_IN1, &db));
    OutBuf<uint32_t> o;
    RH_TRY(o.prepare(ctx, out, (size_t)n, S_OUT0));
    distances_kernel<2><<<cdiv(n, 256), 256, 0, st>>>(da, db, (size_t)n, o.dev);
    RH_LAUNCHED(ctx, "distances_kernel");
    RH_TRY(o.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}

int rh_hamming_edges(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                     const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                     uint32_t *out_edges, size_t edges_cap, uint64_t *out_edge_count) {
    return group_impl<8>(ctx, hashes, has_hash, variants, n_variants, low_conf, n, similarity, RH_MAX_SIMILARITY_256,
                         0, 1, nullptr, out_edge_count, out_edges, edges_cap);
}

int rh_find_groups(rh_ctx *ctx, const uint8_t *hashes, int64_t n, int width_bits, uint32_t max_dist,
                   uint32_t *members, size_t members_cap, uint32_t *group_offsets, size_t groups_cap,
                   size_t *n_groups) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    if (!n_groups || (width_bits != 64 && width_bits != 256) || n < 0 || (n > 0 && !hashes) || !group_offsets ||
        groups_cap < 1)
        return fail(ctx, RH_EINVAL, "rh_find_groups: bad arguments");
    *n_groups = 0;
    group_offsets[0] = 0;
    if (n < 2) return RH_OK;
    if (max_dist > (uint32_t)width_bits) max_dist = (uint32_t)width_bits;
    // Device: the exact adjacency as an unordered list of (i < j) pairs.  The list is sized by a
    // first guess and the search repeated once if it overflowed.
    std::vector<uint32_t> edges;
    size_t cap = (size_t)n * 4 + (1u << 20);
    uint64_t count = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        try {
            edges.assign(cap * 2, 0u);
        } catch (const std::bad_alloc &) {
            return fail(ctx, RH_ENOMEM, "rh_find_groups: edge list");
        }
        int s = width_bits == 256
                    ? group_impl<8>(ctx, hashes, nullptr, nullptr, nullptr, nullptr, n, max_dist, 256, 0, 1, nullptr,
                                    &count, edges.data(), cap)
                    : group_impl<2>(ctx, hashes, nullptr, nullptr, nullptr, nullptr, n, max_dist, 64, 0, 1, nullptr,
                                    &count, edges.data(), cap);
        if (s != RH_OK) return s;
        if (count <= cap) break;
        // the greedy star clustering needs the adjacency itself; inputs with O(n^2) edges (a library of
        // identical files) are what the union-find entry points are for -- they never materialise edges
        if (count > (uint64_t(1) << 30))
            return fail(ctx, RH_EUNSUPPORTED, "rh_find_groups: more than 2^30 edges; use rh_hamming_group (connected components)");
        cap = (size_t)count;
    }
    // Host: CSR adjacency (both directions, ascending), then the sequential greedy star
    // clustering of hamminghash.rs:245-270 -- sequential in the reference too.
    std::vector<uint32_t> deg((size_t)n + 1, 0u);
    for (uint64_t k = 0; k < count; k++) {
        deg[edges[2 * k] + 1]++;
        deg[edges[2 * k + 1] + 1]++;
    }
    for (int64_t i = 0; i < n; i++) deg[i + 1] += deg[i];
    std::vector<uint32_t> adj((size_t)count * 2), fill(deg.begin(), deg.end() - 1);
    for (uint64_t k = 0; k < count; k++) {
        uint32_t a = edges[2 * k], b = edges[2 * k + 1];
        adj[fill[a]++] = b;
        adj[fill[b]++] = a;
    }
    std::vector<uint8_t> visited((size_t)n, 0);
    size_t ng = 0, nm = 0;
    for (int64_t i = 0; i < n; i++) {
        if (visited[i] || deg[i + 1] == deg[i]) continue;
        visited[i] = 1;
        std::sort(adj.begin() + deg[i], adj.begin() + deg[i + 1]);
        size_t len = 1;
        for (uint32_t k = deg[i]; k < deg[i + 1]; k++)
            if (!visited[adj[k]]) len++;
        if (len > 1) {
            if (ng + 1 >= groups_cap || nm + len > members_cap || !members)
                return fail(ctx, RH_EINVAL, "rh_find_groups: output capacity too small");
            members[nm++] = (uint32_t)i;
        }
        for (uint32_t k = deg[i]; k < deg[i + 1]; k++) {
            uint32_t v = adj[k];
            if (!visited[v]) {
                visited[v] = 1;
                if (len > 1) members[nm++] = v;
            }
        }
        if (len > 1) {
            group_offsets[++ng] = (uint32_t)nm;
        }
    }
    *n_groups = ng;
    return RH_OK;
}

int rh_group_max_dist(rh_ctx *ctx, const uint8_t *pivot_variants, const uint8_t *n_pivot_variants,
                      const uint8_t *member_hashes, const uint32_t *member_group, int64_t n_members, int64_t n_groups,
                      uint32_t *out_max_dist) {
    using namespace rh;
    if (!ctx) return RH_EINVAL;
    if (n_members < 0 || n_groups < 0 || (n_groups > 0 && (!pivot_variants || !out_max_dist)) ||
        (n_members > 0 && (!member_hashes || !member_group)))
        return fail(ctx, RH_EINVAL, "rh_group_max_dist: bad arguments");
    if (n_groups == 0) return RH_OK;
    RH_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t *d_piv, *d_nv, *d_mem;
    const uint32_t *d_mg;
    RH_TRY(stage_in(ctx, pivot_variants, (size_t)n_groups * 256, S_IN0, &d_piv));
    RH_TRY(stage_in(ctx, n_pivot_variants, (size_t)n_groups, S_IN1, &d_nv));
    RH_TRY(stage_in(ctx, member_hashes, (size_t)n_members * 32, S_IN2, &d_mem));
    RH_TRY(stage_in(ctx, member_group, (size_t)n_members, S_IN3, &d_mg));
    OutBuf<uint32_t> o;
    RH_TRY(o.prepare(ctx, out_max_dist, (size_t)n_groups, S_OUT0));
    RH_CUDA(ctx, cudaMemsetAsync(o.dev, 0, (size_t)n_groups * 4, st));
    if (n_members > 0) {
        group_max_dist_kernel<<<cdiv(n_members, 256), 256, 0, st>>>(d_piv, d_nv, d_mem, d_mg, (size_t)n_members, o.dev);
        RH_LAUNCHED(ctx, "group_max_dist_kernel");
    }
    RH_TRY(o.finish(ctx));
    RH_CUDA(ctx, cudaStreamSynchronize(st));
    return RH_OK;
}

}  // extern "C"
