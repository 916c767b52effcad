// group_multi_harness.cpp -- a plain C++ program (no Python, no torch at run time) that groups a hash
// file on every GPU of the box through rh_group / rh_hamming_group_multi, the way a single-process
// caller such as the reference's scan thread (scanner.rs:1550-1551) would, and writes the labels.
//
//   group_multi_harness <in.bin> <out.bin> <similarity> <n_gpus (0 = all)> <flags>
//
// in.bin : i64 n, u8 has_variants, u8 has_low_conf, u8 has_has_hash, u8 pad[5], hashes n*32,
//          [variants n*256], [low_conf n], [has_hash n]
// out.bin: u64 edge count, f64 wall ms, f64 slowest tile ms, f64 sum of tile ms, u32 n_gpus, u32 nccl version,
//          labels n*u32 (multi-GPU), labels n*u32 (single GPU, rh_hamming_group on device 0)
// tests/test_gpu_group_multi.py writes the input, runs this binary and compares both label arrays
// with the CPU oracle.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/rupphash_b200.h"

static bool read_all(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s in.bin out.bin similarity n_gpus flags\n", argv[0]);
        return 2;
    }
    const unsigned similarity = (unsigned)atoi(argv[3]);
    const int n_gpus = atoi(argv[4]);
    const unsigned flags = (unsigned)atoi(argv[5]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t n = 0;
    uint8_t hdr[8];
    if (!read_all(f, &n, 8) || !read_all(f, hdr, 8) || n < 0) return 4;
    // page-locked buffers, as a caller that cares about the copy rate would use
    void *p_h = nullptr, *p_v = nullptr, *p_l = nullptr, *p_hh = nullptr;
    if (rh_alloc_pinned((size_t)n * 32, &p_h) != RH_OK) return 5;
    if (hdr[0] && rh_alloc_pinned((size_t)n * 256, &p_v) != RH_OK) return 5;
    if (hdr[1] && rh_alloc_pinned((size_t)n, &p_l) != RH_OK) return 5;
    if (hdr[2] && rh_alloc_pinned((size_t)n, &p_hh) != RH_OK) return 5;
    if (!read_all(f, p_h, (size_t)n * 32) || (p_v && !read_all(f, p_v, (size_t)n * 256)) ||
        (p_l && !read_all(f, p_l, (size_t)n)) || (p_hh && !read_all(f, p_hh, (size_t)n)))
        return 6;
    fclose(f);

    rh_group *g = nullptr;
    int rc = rh_group_create(nullptr, n_gpus, flags, &g);
    if (rc != RH_OK) {
        fprintf(stderr, "rh_group_create: %d\n", rc);
        return 7;
    }
    std::vector<uint32_t> multi((size_t)n), single((size_t)n);
    uint64_t edges = 0, edges1 = 0;
    double times[8] = {};
    for (int rep = 0; rep < 3; rep++) {   // the last repetition is the timed one (buffers warm)
        rc = rh_hamming_group_multi(g, (const uint8_t *)p_h, (const uint8_t *)p_hh, (const uint8_t *)p_v, nullptr,
                                    (const uint8_t *)p_l, n, similarity, multi.data(), &edges);
        if (rc != RH_OK) {
            fprintf(stderr, "rh_hamming_group_multi: %d %s\n", rc, rh_group_last_error(g));
            return 8;
        }
    }
    rh_group_last_times(g, times, 8);
    rh_ctx *c0 = rh_group_ctx(g, 0);
    rc = rh_hamming_group(c0, (const uint8_t *)p_h, (const uint8_t *)p_hh, (const uint8_t *)p_v, nullptr, (const uint8_t *)p_l, n,
                          similarity, single.data(), &edges1);
    if (rc != RH_OK) {
        fprintf(stderr, "rh_hamming_group: %d %s\n", rc, rh_last_error(c0));
        return 9;
    }
    if (edges1 != edges) {
        fprintf(stderr, "edge count: %llu on the group, %llu on one GPU\n", (unsigned long long)edges, (unsigned long long)edges1);
        return 10;
    }
    int nccl = 0, steal = 0;
    rh_group_info(g, &nccl, &steal);
    const uint32_t ng = (uint32_t)rh_group_size(g), nv = (uint32_t)nccl;
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 11;
    fwrite(&edges, 8, 1, o);
    fwrite(&times[0], 8, 1, o);
    fwrite(&times[1], 8, 1, o);
    fwrite(&times[3], 8, 1, o);
    fwrite(&ng, 4, 1, o);
    fwrite(&nv, 4, 1, o);
    fwrite(multi.data(), 4, (size_t)n, o);
    fwrite(single.data(), 4, (size_t)n, o);
    fclose(o);
    printf("group harness ok: n=%lld gpus=%u nccl=%d stealing=%d edges=%llu wall=%.3f ms tile(max)=%.3f ms\n", (long long)n, ng,
           nccl, steal, (unsigned long long)edges, times[0], times[1]);
    rh_group_destroy(g);
    rh_free_pinned(p_h);
    if (p_v) rh_free_pinned(p_v);
    if (p_l) rh_free_pinned(p_l);
    if (p_hh) rh_free_pinned(p_hh);
    return 0;
}
