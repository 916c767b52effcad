// rupphash.hpp -- header-only C++17 mirror of the reference's Rust interface for the two hot
// paths, forwarding to the C ABI (include/rupphash_b200.h).  Names, argument meaning and error
// behaviour follow the reference (file:line cited per function); Option<T> becomes
// std::optional<T>, panics become exceptions thrown on the HOST side of the ABI only.
#pragma once
#include <array>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <tuple>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/rupphash_b200.h"

namespace rupphash {

class Context {
public:
    explicit Context(int device = 0) {
        if (rh_ctx_create(device, &ctx_) != RH_OK)
            throw std::runtime_error("rh_ctx_create failed: no usable CUDA device (no CPU fallback)");
    }
    ~Context() { rh_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    rh_ctx *get() const { return ctx_; }
    void check(int rc) const {
        if (rc == RH_OK) return;
        std::string msg = rh_last_error(ctx_);
        if (rc == RH_EINVAL) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

private:
    rh_ctx *ctx_ = nullptr;
};

// image::DynamicImage after decode: interleaved u8 pixels (pdqhash.rs:268-284 accepts these three)
struct ImageView {
    const uint8_t *pixels;
    int width, height;
    rh_layout layout;
};

namespace pdqhash {

using Hash = std::array<uint8_t, RH_PDQ_HASH_BYTES>;

// pdqhash.rs:48-51
struct PdqFeatures {
    std::array<float, RH_PDQ_COEFFS> coefficients;
    // pdqhash.rs:59-61
    Hash to_hash(const Context &c) const {
        Hash h;
        c.check(rh_pdq_hash_from_coeffs(c.get(), coefficients.data(), 1, h.data()));
        return h;
    }
    // pdqhash.rs:71-87
    std::array<Hash, 8> generate_dihedral_hashes(const Context &c) const {
        std::array<Hash, 8> out;
        c.check(rh_pdq_dihedral_from_coeffs(c.get(), coefficients.data(), 1, out[0].data()));
        return out;
    }
};

// pdqhash.rs:166-196: nullopt when width or height < 5
inline std::optional<std::pair<PdqFeatures, float>> generate_pdq_features(const Context &c, const ImageView &img) {
    PdqFeatures f;
    float q = 0.f;
    uint8_t valid = 0;
    c.check(rh_pdq_hash_batch(c.get(), img.pixels, img.layout, 1, img.width, img.height, 0, 0, nullptr, &q,
                              f.coefficients.data(), nullptr, &valid));
    if (!valid) return std::nullopt;
    return std::make_pair(f, q);
}

// pdqhash.rs:199-201
inline std::optional<std::pair<Hash, float>> generate_pdq(const Context &c, const ImageView &img) {
    Hash h;
    float q = 0.f;
    uint8_t valid = 0;
    c.check(rh_pdq_hash_batch(c.get(), img.pixels, img.layout, 1, img.width, img.height, 0, 0, h.data(), &q, nullptr,
                              nullptr, &valid));
    if (!valid) return std::nullopt;
    return std::make_pair(h, q);
}

// The scanner's batched form (scanner.rs:1409-1412 over a batch of same-sized decoded images): n
// interleaved images back to back in host or device memory.  valid[i] == 0 is the reference's None.
struct BatchResult {
    std::vector<Hash> hashes;
    std::vector<float> quality;
    std::vector<uint8_t> valid;
    std::vector<std::array<float, RH_PDQ_COEFFS>> coefficients;   // empty unless want_coeffs
    std::vector<std::array<Hash, 8>> dihedral;                    // empty unless want_dihedral
};
inline BatchResult hash_batch(const Context &c, const uint8_t *pixels, rh_layout layout, int64_t n, int width, int height,
                              bool want_coeffs = false, bool want_dihedral = false) {
    BatchResult r;
    r.hashes.resize((size_t)n);
    r.quality.resize((size_t)n);
    r.valid.resize((size_t)n);
    if (want_coeffs) r.coefficients.resize((size_t)n);
    if (want_dihedral) r.dihedral.resize((size_t)n);
    if (n == 0) return r;
    c.check(rh_pdq_hash_batch(c.get(), pixels, layout, n, width, height, 0, 0, r.hashes[0].data(), r.quality.data(),
                              want_coeffs ? r.coefficients[0].data() : nullptr,
                              want_dihedral ? r.dihedral[0][0].data() : nullptr, r.valid.data()));
    return r;
}

}  // namespace pdqhash

namespace phash {
// phash.rs:150-255
inline uint64_t rotate_hash_90(uint64_t h) { return rh_phash_rotate_90(h); }
inline uint64_t rotate_hash_180(uint64_t h) { return rh_phash_rotate_180(h); }
inline uint64_t rotate_hash_270(uint64_t h) { return rh_phash_rotate_270(h); }
inline uint64_t flip_hash_horizontal(uint64_t h) { return rh_phash_flip_horizontal(h); }
inline std::vector<uint64_t> generate_dihedral_hashes(uint64_t h) {
    std::vector<uint64_t> v(8);
    rh_phash_dihedral(h, v.data());
    return v;
}
inline uint64_t calculate_rotation_invariant_hash(uint64_t h) { return rh_phash_rotation_invariant(h); }
// DctPhash::hash_image (phash.rs:48-83): 32x32 Triangle resize, Rec.709 luma, 32x32 DCT, median of the
// low 8x8 block -> 64 bits.  Images narrower or lower than 1 px are rejected by the library.
struct DctPhash {
    static uint64_t hash_image(const Context &c, const ImageView &img) {
        uint64_t h = 0;
        c.check(rh_phash_batch(c.get(), img.pixels, img.layout, 1, img.width, img.height, 0, 0, &h, nullptr));
        return h;
    }
};
}  // namespace phash

namespace hamminghash {

constexpr uint32_t MAX_SIMILARITY_64 = RH_MAX_SIMILARITY_64;    // hamminghash.rs:5
constexpr uint32_t MAX_SIMILARITY_256 = RH_MAX_SIMILARITY_256;  // hamminghash.rs:8

// hamminghash.rs:55-58 (batch of 1)
inline uint32_t hamming_distance(const Context &c, const pdqhash::Hash &a, const pdqhash::Hash &b) {
    uint32_t d = 0;
    c.check(rh_hamming_distances(c.get(), a.data(), b.data(), 1, &d));
    return d;
}

// hamminghash.rs:34-36 (batch of 1)
inline uint32_t hamming_distance(const Context &c, uint64_t a, uint64_t b) {
    uint32_t d = 0;
    c.check(rh_hamming_distances_u64(c.get(), &a, &b, 1, &d));
    return d;
}

// hamminghash.rs:82-149: on the device the search is all-pairs, the index only holds the hashes
struct MIHIndex {
    std::vector<pdqhash::Hash> db_hashes;
    static MIHIndex make(std::vector<pdqhash::Hash> hashes) { return MIHIndex{std::move(hashes)}; }
    size_t len() const { return db_hashes.size(); }
    const pdqhash::Hash &hash(uint32_t dense_id) const { return db_hashes[dense_id]; }
};

// hamminghash.rs:191-271
inline std::vector<std::vector<uint32_t>> find_groups(const Context &c, const MIHIndex &index, uint32_t max_dist) {
    const size_t n = index.len();
    std::vector<uint32_t> members(n ? n : 1), offsets(n / 2 + 2);
    size_t ng = 0;
    c.check(rh_find_groups(c.get(), n ? index.db_hashes[0].data() : nullptr, (int64_t)n, 256, max_dist, members.data(),
                           members.size(), offsets.data(), offsets.size(), &ng));
    std::vector<std::vector<uint32_t>> groups(ng);
    for (size_t g = 0; g < ng; g++) groups[g].assign(members.begin() + offsets[g], members.begin() + offsets[g + 1]);
    return groups;
}

}  // namespace hamminghash

namespace scanner {

// scanner.rs:1416-1418
inline uint16_t quality_100(float q) {
    float v = q * 100.0f;
    v = v >= 0.f ? (float)(long)(v + 0.5f) : -(float)(long)(-v + 0.5f);
    if (v < 0.f) v = 0.f;
    if (v > 100.f) v = 100.f;
    return (uint16_t)v;
}

// scanner.rs:1588-1594, :1631-1636: unknown quality counts as good
inline bool is_low_confidence(std::optional<uint16_t> quality100) { return quality100 && *quality100 < 50; }

struct GroupResult {
    std::vector<std::vector<uint32_t>> groups;  // members ascending, groups ordered by first member
    uint64_t comparison_count;                  // scanner.rs:1778
};

namespace detail {
inline GroupResult groups_from_labels(const std::vector<uint32_t> &label, size_t n, uint64_t count) {
    std::vector<uint32_t> size(n, 0), slot(n, UINT32_MAX);
    for (size_t i = 0; i < n; i++) size[label[i]]++;
    GroupResult r;
    r.comparison_count = count;
    for (size_t i = 0; i < n; i++) {
        const uint32_t root = label[i];
        if (size[root] < 2) continue;
        if (slot[root] == UINT32_MAX) {
            slot[root] = (uint32_t)r.groups.size();
            r.groups.emplace_back();
        }
        r.groups[slot[root]].push_back((uint32_t)i);
    }
    return r;
}
}  // namespace detail
using detail::groups_from_labels;

// scanner.rs:1640-1817.  variants: n x 8 hashes or empty (= every file queries with its own hash);
// low_conf / has_hash: n flags or empty.  similarity > 63 throws (the reference asserts, :1650-1655).
inline GroupResult group_files_generic(const Context &c, const std::vector<pdqhash::Hash> &hashes, uint32_t similarity,
                                       const std::vector<uint8_t> &has_hash = {},
                                       const std::vector<std::array<pdqhash::Hash, 8>> &variants = {},
                                       const std::vector<uint8_t> &low_conf = {}) {
    const size_t n = hashes.size();
    std::vector<uint32_t> label(n ? n : 1);
    uint64_t count = 0;
    c.check(rh_hamming_group(c.get(), n ? hashes[0].data() : nullptr, has_hash.empty() ? nullptr : has_hash.data(),
                             variants.empty() ? nullptr : variants[0][0].data(), nullptr,
                             low_conf.empty() ? nullptr : low_conf.data(), (int64_t)n, similarity, label.data(), &count));
    return groups_from_labels(label, n, count);   // groups_map (scanner.rs:1809-1817) in canonical form
}


// The same search over u64 hashes (impl HammingHash for u64, hamminghash.rs:23-41).  NO reference caller groups
// u64 hashes (phash.rs is reached only from the phash_test demo); similarity <= 15 is what the
// reference's MIH could serve for them (hamminghash.rs:5), the exact search accepts up to 63.
inline GroupResult group_files_generic_u64(const Context &c, const std::vector<uint64_t> &hashes, uint32_t similarity,
                                           const std::vector<uint8_t> &has_hash = {},
                                           const std::vector<std::array<uint64_t, 8>> &variants = {}) {
    const size_t n = hashes.size();
    std::vector<uint32_t> label(n ? n : 1);
    uint64_t count = 0;
    c.check(rh_hamming_group_u64(c.get(), n ? hashes.data() : nullptr, has_hash.empty() ? nullptr : has_hash.data(),
                                 variants.empty() ? nullptr : variants[0].data(), nullptr, nullptr, (int64_t)n, similarity,
                                 label.data(), &count));
    return detail::groups_from_labels(label, n, count);
}

// One rank's share of group_files_generic (tiles t with t % world == rank): its union-find forest and
// edge count.  The forests of all ranks go to merge_forests; the edge counts add up.
inline std::pair<std::vector<uint32_t>, uint64_t> group_files_shard(const Context &c, const std::vector<pdqhash::Hash> &hashes,
                                                                    uint32_t similarity, int rank, int world,
                                                                    const std::vector<uint8_t> &low_conf = {}) {
    const size_t n = hashes.size();
    std::vector<uint32_t> parent(n ? n : 1);
    uint64_t count = 0;
    c.check(rh_hamming_group_shard(c.get(), n ? hashes[0].data() : nullptr, nullptr, nullptr, nullptr,
                                   low_conf.empty() ? nullptr : low_conf.data(), (int64_t)n, similarity, rank, world,
                                   parent.data(), &count));
    parent.resize(n);
    return {std::move(parent), count};
}
// forests: world x n, rank-major (what an all-gather of the shard forests produces)
inline GroupResult merge_forests(const Context &c, const std::vector<uint32_t> &forests, int world, uint64_t edge_count) {
    const size_t n = world ? forests.size() / (size_t)world : 0;
    std::vector<uint32_t> label(n ? n : 1);
    c.check(rh_uf_merge(c.get(), forests.data(), world, (int64_t)n, label.data()));
    return detail::groups_from_labels(label, n, edge_count);
}

// max_dist of analyze_group_with_features (scanner.rs:2217-2241) for many groups at once: per group, the
// maximum over its members of the minimum distance to the pivot's dihedral variants.
inline std::vector<uint32_t> group_max_dist(const Context &c, const std::vector<std::array<pdqhash::Hash, 8>> &pivot_variants,
                                            const std::vector<pdqhash::Hash> &member_hashes,
                                            const std::vector<uint32_t> &member_group) {
    std::vector<uint32_t> out(pivot_variants.size());
    if (pivot_variants.empty()) return out;
    c.check(rh_group_max_dist(c.get(), pivot_variants[0][0].data(), nullptr,
                              member_hashes.empty() ? nullptr : member_hashes[0].data(), member_group.data(),
                              (int64_t)member_hashes.size(), (int64_t)pivot_variants.size(), out.data()));
    return out;
}

// ---------------------------------------------------------------------------------------------------
// The hash loop of scan_and_group (scanner.rs:1202-1521) restructured into batches, natively: the caller's
// decode workers (the reference's rayon pool, scanner.rs:1188-1205) call push() concurrently with decoded
// images of any size; same-sized images are copied into page-locked staging batches (capped by count and
// by bytes, at most `max_open_shapes` partly filled batches); ONE submitter thread owns the rh_ctx and
// keeps `inflight` batches queued with rh_pdq_hash_batch_async, handing every result to `on_result`
// (called on the submitter thread: the DbUpdate fan-out of scanner.rs:1495-1518).  finish() flushes the
// partial batches and joins.  Images smaller than 5 x 5 are reported with valid = false (pdqhash.rs:167-169).
struct FileHash {
    size_t index;       // the caller's file index
    bool valid;         // false = the reference's None
    pdqhash::Hash hash;
    float quality;
    uint16_t quality_100;
    std::array<float, RH_PDQ_COEFFS> coefficients;   // filled when want_coeffs
};

class BatchFeeder {
public:
    using ResultFn = std::function<void(const FileHash &)>;
    BatchFeeder(Context &ctx, ResultFn on_result, bool want_coeffs = true, size_t batch_images = 256,
                size_t batch_bytes = size_t(256) << 20, size_t max_open_shapes = 8, size_t inflight = 2)
        : ctx_(ctx), on_result_(std::move(on_result)), want_coeffs_(want_coeffs), batch_images_(batch_images),
          batch_bytes_(batch_bytes), max_open_(max_open_shapes), inflight_(inflight ? inflight : 1) {
        submitter_ = std::thread([this] { run(); });
    }
    ~BatchFeeder() {
        try {
            finish();
        } catch (...) {
        }
        for (auto &kv : free_)
            for (Batch *b : kv.second) release(b);
    }
    BatchFeeder(const BatchFeeder &) = delete;
    BatchFeeder &operator=(const BatchFeeder &) = delete;

    // thread-safe; copies the pixels (rows must be tight: width * channels bytes)
    void push(size_t index, const ImageView &img) {
        const int ch = img.layout == RH_LAYOUT_RGB8 ? 3 : (img.layout == RH_LAYOUT_RGBA8 ? 4 : 1);
        if (img.width < 5 || img.height < 5) {
            FileHash r{};
            r.index = index;
            std::lock_guard<std::mutex> lk(small_m_);
            small_.push_back(r);
            return;
        }
        const size_t bytes = (size_t)img.width * img.height * ch;
        const Key key{img.width, img.height, (int)img.layout};
        Batch *b = nullptr, *evicted = nullptr;
        size_t slot = 0;
        {
            std::unique_lock<std::mutex> lk(m_);
            if (failed_) throw std::runtime_error(error_);
            auto it = open_.find(key);
            if (it == open_.end()) {
                if (open_.size() >= max_open_) {   // send the least recently used partial batch early
                    auto lru = open_.begin();
                    for (auto j = open_.begin(); j != open_.end(); ++j)
                        if (j->second->stamp < lru->second->stamp) lru = j;
                    evicted = lru->second;
                    open_.erase(lru);
                }
                size_t cap = batch_bytes_ / bytes;
                if (cap < 1) cap = 1;
                if (cap > batch_images_) cap = batch_images_;
                b = acquire(key, cap, bytes);
                open_[key] = b;
            } else
                b = it->second;
            slot = b->idx.size();
            b->idx.push_back(index);
            b->stamp = ++stamp_;
            if (b->idx.size() >= b->cap) open_.erase(key);
        }
        std::memcpy(b->pixels + slot * bytes, img.pixels, bytes);   // outside the lock: the workers copy in parallel
        bool full;
        {
            std::lock_guard<std::mutex> lk(m_);
            b->filled++;
            full = b->idx.size() >= b->cap && b->filled == b->idx.size();
            cv_fill_.notify_all();
        }
        if (evicted) enqueue(evicted);
        if (full) enqueue(b);
    }

    // flush the partial batches, wait for every result; rethrows a device error
    void finish() {
        if (!submitter_.joinable()) return;
        std::vector<Batch *> rest;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (auto &kv : open_) rest.push_back(kv.second);
            open_.clear();
        }
        for (Batch *b : rest) enqueue(b);
        {
            std::lock_guard<std::mutex> lk(m_);
            done_ = true;
            cv_work_.notify_all();
        }
        submitter_.join();
        for (const FileHash &r : small_) on_result_(r);
        small_.clear();
        if (failed_) throw std::runtime_error(error_);
    }

private:
    struct Key {
        int w, h, layout;
        bool operator<(const Key &o) const { return std::tie(w, h, layout) < std::tie(o.w, o.h, o.layout); }
    };
    struct Batch {
        Key key;
        size_t cap = 0, bytes = 0, filled = 0;
        uint64_t stamp = 0, ticket = 0;
        uint8_t *pixels = nullptr;   // page-locked: cap images
        uint8_t *hash = nullptr, *valid = nullptr;
        float *quality = nullptr, *coeffs = nullptr;
        std::vector<size_t> idx;
    };
    Batch *acquire(const Key &key, size_t cap, size_t bytes) {   // under m_
        auto &pool = free_[std::make_pair(key, cap)];
        if (!pool.empty()) {
            Batch *b = pool.back();
            pool.pop_back();
            b->idx.clear();
            b->filled = 0;
            return b;
        }
        Batch *b = new Batch();
        b->key = key;
        b->cap = cap;
        b->bytes = bytes;
        void *p = nullptr;
        const size_t res = cap * (32 + 1 + 4 + (want_coeffs_ ? RH_PDQ_COEFFS * 4 : 0));
        if (rh_alloc_pinned(cap * bytes + res + 64, &p) != RH_OK) {
            delete b;
            throw std::bad_alloc();
        }
        b->pixels = (uint8_t *)p;
        uint8_t *q = b->pixels + ((cap * bytes + 15) & ~size_t(15));
        b->quality = (float *)q;
        q += cap * 4;
        if (want_coeffs_) {
            b->coeffs = (float *)q;
            q += cap * RH_PDQ_COEFFS * 4;
        }
        b->hash = q;
        q += cap * 32;
        b->valid = q;
        return b;
    }
    void release(Batch *b) {
        rh_free_pinned(b->pixels);
        delete b;
    }
    void enqueue(Batch *b) {
        std::unique_lock<std::mutex> lk(m_);
        cv_fill_.wait(lk, [&] { return b->filled == b->idx.size(); });   // copies still in progress
        if (b->idx.empty()) {
            free_[std::make_pair(b->key, b->cap)].push_back(b);
            return;
        }
        cv_work_.wait(lk, [&] { return work_.size() < 2 * inflight_ || failed_; });   // back-pressure on the decoders
        work_.push_back(b);
        cv_work_.notify_all();
    }
    void publish(Batch *b) {
        for (size_t k = 0; k < b->idx.size(); k++) {
            FileHash r{};
            r.index = b->idx[k];
            r.valid = b->valid[k] != 0;
            if (r.valid) {
                std::memcpy(r.hash.data(), b->hash + 32 * k, 32);
                r.quality = b->quality[k];
                r.quality_100 = quality_100(r.quality);
                if (want_coeffs_) std::memcpy(r.coefficients.data(), b->coeffs + RH_PDQ_COEFFS * k, RH_PDQ_COEFFS * 4);
            }
            on_result_(r);
        }
        std::lock_guard<std::mutex> lk(m_);
        free_[std::make_pair(b->key, b->cap)].push_back(b);
    }
    void run() {
        std::deque<Batch *> pending;
        for (;;) {
            Batch *b = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return !work_.empty() || done_; });
                if (work_.empty()) break;
                b = work_.front();
                work_.pop_front();
                cv_work_.notify_all();
            }
            if (failed_) continue;
            int rc = rh_pdq_hash_batch_async(ctx_.get(), b->pixels, b->key.layout, (int64_t)b->idx.size(), b->key.w, b->key.h, 0, 0,
                                             b->hash, b->quality, b->coeffs, nullptr, b->valid, &b->ticket);
            if (rc != RH_OK) {
                fail(rh_last_error(ctx_.get()));
                continue;
            }
            pending.push_back(b);
            if (pending.size() >= inflight_) drain_one(pending);
        }
        while (!pending.empty() && !failed_) drain_one(pending);
        if (failed_) rh_ctx_sync(ctx_.get());
    }
    void drain_one(std::deque<Batch *> &pending) {
        Batch *b = pending.front();
        pending.pop_front();
        if (rh_ctx_wait(ctx_.get(), b->ticket) != RH_OK) {
            fail(rh_last_error(ctx_.get()));
            return;
        }
        publish(b);
    }
    void fail(const char *msg) {
        std::lock_guard<std::mutex> lk(m_);
        failed_ = true;
        error_ = msg;
        cv_work_.notify_all();
    }

    Context &ctx_;
    ResultFn on_result_;
    bool want_coeffs_;
    size_t batch_images_, batch_bytes_, max_open_, inflight_;
    std::mutex m_, small_m_;
    std::condition_variable cv_work_, cv_fill_;
    std::map<Key, Batch *> open_;
    std::map<std::pair<Key, size_t>, std::vector<Batch *>> free_;
    std::deque<Batch *> work_;
    std::vector<FileHash> small_;
    uint64_t stamp_ = 0;
    bool done_ = false, failed_ = false;
    std::string error_;
    std::thread submitter_;
};

// Every GPU of the box from this one process (rh_group): what group_with_pdqhash (scanner.rs:1827-1832) calls.
class Group {
public:
    explicit Group(int n_dev = 0, unsigned flags = 0) {
        if (rh_group_create(nullptr, n_dev, flags, &g_) != RH_OK)
            throw std::runtime_error("rh_group_create failed (no CUDA devices, no peer access, or NCCL unavailable)");
    }
    ~Group() { rh_group_destroy(g_); }
    Group(const Group &) = delete;
    Group &operator=(const Group &) = delete;
    int size() const { return rh_group_size(g_); }
    GroupResult group_files_generic(const std::vector<pdqhash::Hash> &hashes, uint32_t similarity,
                                    const std::vector<uint8_t> &has_hash = {},
                                    const std::vector<std::array<pdqhash::Hash, 8>> &variants = {},
                                    const std::vector<uint8_t> &low_conf = {}) const {
        const size_t n = hashes.size();
        std::vector<uint32_t> label(n ? n : 1);
        uint64_t count = 0;
        int rc = rh_hamming_group_multi(g_, n ? hashes[0].data() : nullptr, has_hash.empty() ? nullptr : has_hash.data(),
                                        variants.empty() ? nullptr : variants[0][0].data(), nullptr,
                                        low_conf.empty() ? nullptr : low_conf.data(), (int64_t)n, similarity, label.data(), &count);
        if (rc == RH_EINVAL) throw std::invalid_argument(rh_group_last_error(g_));
        if (rc != RH_OK) throw std::runtime_error(rh_group_last_error(g_));
        return detail::groups_from_labels(label, n, count);
    }

private:
    rh_group *g_ = nullptr;
};

}  // namespace scanner
}  // namespace rupphash
