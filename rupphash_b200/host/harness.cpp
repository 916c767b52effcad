// harness.cpp -- drives the C++ mirror the way the reference's own unit tests drive the Rust
// functions (hamminghash.rs:283-332 KAT, pdqhash.rs:641-647-style None case).  Needs a GPU to run;
// tests/test_host_cpp.py compiles it everywhere and runs it under -m gpu.
#include <cstdio>
#include <cstring>

#include "rupphash.hpp"

using namespace rupphash;

int main() {
    Context ctx(0);
    // hamminghash.rs:305-331: a PDQ hash and one 30 bits away group together at similarity 30
    std::vector<pdqhash::Hash> hashes(3);
    for (auto &h : hashes) h.fill(0);
    hashes[1][0] = hashes[1][1] = hashes[1][2] = 0xFF;
    hashes[1][3] = 0x3F;
    hashes[2].fill(0xAA);
    auto r = scanner::group_files_generic(ctx, hashes, 30);
    if (r.groups.size() != 1 || r.groups[0] != std::vector<uint32_t>{0, 1} || r.comparison_count != 1) return 1;
    if (hamminghash::hamming_distance(ctx, hashes[0], hashes[1]) != 30) return 2;
    auto star = hamminghash::find_groups(ctx, hamminghash::MIHIndex::make(hashes), 30);
    if (star.size() != 1 || star[0].size() != 2) return 3;
    bool threw = false;
    try {
        scanner::group_files_generic(ctx, hashes, 64);
    } catch (const std::invalid_argument &) {
        threw = true;
    }
    if (!threw) return 4;
    // pdqhash.rs:167-169: images narrower than 5 px hash to None
    std::vector<uint8_t> px(4 * 100 * 3, 7);
    if (pdqhash::generate_pdq_features(ctx, ImageView{px.data(), 4, 100, RH_LAYOUT_RGB8}).has_value()) return 5;
    std::vector<uint8_t> img(512 * 384 * 3);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 24);
    auto f = pdqhash::generate_pdq_features(ctx, ImageView{img.data(), 512, 384, RH_LAYOUT_RGB8});
    auto g = pdqhash::generate_pdq(ctx, ImageView{img.data(), 512, 384, RH_LAYOUT_RGB8});
    if (!f || !g) return 6;
    if (f->first.to_hash(ctx) != g->first) return 7;
    if (f->first.generate_dihedral_hashes(ctx)[0] != g->first) return 8;
    if (scanner::quality_100(0.495f) != 50 || scanner::quality_100(1.0f) != 100) return 9;
    if (phash::generate_dihedral_hashes(0x0123456789ABCDEFull)[0] != 0x0123456789ABCDEFull) return 10;
    std::puts("harness ok");
    return 0;
}
