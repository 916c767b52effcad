// rupphash.hpp -- header-only C++17 mirror of the reference's Rust interface for the two hot
// paths, forwarding to the C ABI (include/rupphash_b200.h).  Names, argument meaning and error
// behaviour follow the reference (file:line cited per function); Option<T> becomes
// std::optional<T>, panics become exceptions thrown on the HOST side of the ABI only.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
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

struct GroupResult {
    std::vector<std::vector<uint32_t>> groups;  // members ascending, groups ordered by first member
    uint64_t comparison_count;                  // scanner.rs:1778
};

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
    // groups_map (scanner.rs:1809-1817) in canonical form
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

}  // namespace scanner
}  // namespace rupphash
