// harness.cpp -- drives the C++ mirror the way the reference's own unit tests drive the Rust
// functions (hamminghash.rs:283-332 KAT, pdqhash.rs:641-647-style None case).  Needs a GPU to run;
// tests/test_host_cpp.py compiles it everywhere and runs it under -m gpu.
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "rupphash.hpp"

using namespace rupphash;

// `harness feeder-bench [images] [threads]`: throughput of the native feeder on 1024 x 768 RGB8 images that the
// "decode" threads copy out of pageable memory (the cost of handing a decoded image over), results per file
static int feeder_bench(int n_images, int n_threads) {
    Context ctx(0);
    const int w = 1024, h = 768;
    const size_t bytes = (size_t)w * h * 3;
    std::vector<std::vector<uint8_t>> pool(64, std::vector<uint8_t>(bytes));
    uint32_t lcg = 99;
    for (auto &img : pool)
        for (size_t i = 0; i < bytes; i += 4) {
            lcg = lcg * 1664525u + 1013904223u;
            std::memcpy(&img[i], &lcg, 4);
        }
    // one feeder for the whole run (as a scan would keep it); the first third page-locks the staging batches and
    // is not timed: the clock runs from the n-th to the 3n-th result
    size_t got = 0;
    std::chrono::steady_clock::time_point t1, t2;
    {
        scanner::BatchFeeder feeder(ctx, [&](const scanner::FileHash &r) {
            got += r.valid;
            if (got == (size_t)n_images) t1 = std::chrono::steady_clock::now();
            if (got == (size_t)3 * n_images) t2 = std::chrono::steady_clock::now();
        }, false, 256);
        std::vector<std::thread> workers;
        for (int t = 0; t < n_threads; t++)
            workers.emplace_back([&, t] {
                for (int i = t; i < 3 * n_images; i += n_threads)
                    feeder.push((size_t)i, ImageView{pool[i % 64].data(), w, h, RH_LAYOUT_RGB8});
            });
        for (auto &wk : workers) wk.join();
        feeder.finish();
    }
    if (got != (size_t)3 * n_images) return 30;
    const double best = 2.0 * n_images / std::chrono::duration<double>(t2 - t1).count();
    std::printf("{\"feeder_images_per_s\": %.1f, \"images\": %d, \"decode_threads\": %d, \"note\": \"native BatchFeeder, pageable images -> pinned staging -> async batches, steady state\"}\n",
                best, n_images, n_threads);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "feeder-bench")
        return feeder_bench(argc > 2 ? std::atoi(argv[2]) : 4096, argc > 3 ? std::atoi(argv[3]) : 8);
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
    // batched hashing: the batch of one above is element 1 of a batch of three, bit for bit
    std::vector<uint8_t> batch(3 * img.size());
    for (size_t k = 0; k < 3; k++)
        for (size_t i = 0; i < img.size(); i++) batch[k * img.size() + i] = k == 1 ? img[i] : (uint8_t)(img[i] ^ (uint8_t)(17 * k + 3));
    auto b = pdqhash::hash_batch(ctx, batch.data(), RH_LAYOUT_RGB8, 3, 512, 384, true, true);
    if (!b.valid[0] || !b.valid[1] || !b.valid[2]) return 11;
    if (b.hashes[1] != g->first || b.quality[1] != g->second || b.coefficients[1] != f->first.coefficients) return 12;
    if (b.dihedral[1] != f->first.generate_dihedral_hashes(ctx)) return 13;
    // sharded grouping: two ranks' forests merge into the single-GPU result
    std::vector<pdqhash::Hash> many(3000);
    uint32_t lcg = 12345;
    for (size_t i = 0; i < many.size(); i++)
        for (auto &byte : many[i]) byte = (uint8_t)((lcg = lcg * 1664525u + 1013904223u) >> 24);
    for (size_t i = 0; i < 300; i++) {   // planted near-duplicates: copy + 3 flipped bits
        many[2000 + i] = many[i];
        many[2000 + i][i % 32] ^= 0x15;
    }
    auto single = scanner::group_files_generic(ctx, many, 31);
    auto s0 = scanner::group_files_shard(ctx, many, 31, 0, 2), s1 = scanner::group_files_shard(ctx, many, 31, 1, 2);
    std::vector<uint32_t> forests(s0.first);
    forests.insert(forests.end(), s1.first.begin(), s1.first.end());
    auto merged = scanner::merge_forests(ctx, forests, 2, s0.second + s1.second);
    if (single.groups.size() != 300 || merged.groups != single.groups || merged.comparison_count != single.comparison_count) return 14;
    // max_dist of a group against its pivot's variants (scanner.rs:2217-2241)
    std::vector<std::array<pdqhash::Hash, 8>> pivots(1);
    for (auto &v : pivots[0]) v = many[0];
    auto md = scanner::group_max_dist(ctx, pivots, {many[0], many[2000]}, {0, 0});
    if (md.size() != 1 || md[0] != 3) return 15;
    // 64-bit pHash path: distances, grouping, the image hash of a flat image (all AC terms 0)
    if (hamminghash::hamming_distance(ctx, 0xFFull, 0x0Full) != 4) return 16;
    auto g64 = scanner::group_files_generic_u64(ctx, {0x0ull, 0x7ull, 0xFFFFFFFFFFFFFFFFull}, 3);
    if (g64.groups.size() != 1 || g64.groups[0] != std::vector<uint32_t>{0, 1}) return 17;
    (void)phash::DctPhash::hash_image(ctx, ImageView{img.data(), 512, 384, RH_LAYOUT_RGB8});
    if (scanner::is_low_confidence(std::nullopt) || !scanner::is_low_confidence(49) || scanner::is_low_confidence(50)) return 18;
    // the native feeder: four "decode" threads push images of three shapes (one of them un-hashable), results
    // arrive per file on the submitter thread and equal the direct calls; then the same through every GPU
    {
        std::vector<std::vector<uint8_t>> pix(3);
        const int ws[3] = {512, 256, 4}, hs[3] = {384, 200, 100};
        for (int s = 0; s < 3; s++) {
            pix[s].resize((size_t)ws[s] * hs[s] * 3);
            for (size_t i = 0; i < pix[s].size(); i++) pix[s][i] = (uint8_t)(((i + 7 * s) * 2654435761u) >> 24);
        }
        const size_t n_files = 301;
        std::vector<scanner::FileHash> got(n_files);
        std::vector<int> seen(n_files, 0);
        {
            scanner::BatchFeeder feeder(ctx, [&](const scanner::FileHash &r) { got[r.index] = r; seen[r.index]++; }, true, 32,
                                        size_t(64) << 20, 2, 2);
            std::vector<std::thread> workers;
            for (int t = 0; t < 4; t++)
                workers.emplace_back([&, t] {
                    for (size_t i = t; i < n_files; i += 4) {
                        const int s = (int)(i % 3);
                        feeder.push(i, ImageView{pix[s].data(), ws[s], hs[s], RH_LAYOUT_RGB8});
                    }
                });
            for (auto &w : workers) w.join();
            feeder.finish();
        }
        for (size_t i = 0; i < n_files; i++) {
            if (seen[i] != 1) return 19;
            const int s = (int)(i % 3);
            if (s == 2) {
                if (got[i].valid) return 20;
                continue;
            }
            auto direct = pdqhash::generate_pdq_features(ctx, ImageView{pix[s].data(), ws[s], hs[s], RH_LAYOUT_RGB8});
            if (!direct || !got[i].valid || got[i].quality != direct->second || got[i].coefficients != direct->first.coefficients ||
                got[i].hash != direct->first.to_hash(ctx) || got[i].quality_100 != scanner::quality_100(direct->second))
                return 21;
            if (i > 12) i += 37;   // (spot checks after the first few)
        }
        scanner::Group all;
        auto multi = all.group_files_generic(many, 31);
        if (multi.groups != single.groups || multi.comparison_count != single.comparison_count) return 22;
        std::printf("feeder + group ok on %d GPU(s)\n", all.size());
    }
    std::puts("harness ok");
    return 0;
}
