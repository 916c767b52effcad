// hamming_internal.cuh -- the enqueue-only view of the Hamming search that the multi-GPU group
// (group.cu) drives: everything is queued on the ctx stream, nothing waits for the host.
#pragma once
#include "common.cuh"

namespace rh {

struct HammingPlan {
    int rank = 0, world = 1;                         // static ownership: tile t belongs to rank t mod world
    unsigned long long *shared_counter = nullptr;    // non-null: every GPU claims tiles from this one counter
    cudaEvent_t before_tiles = nullptr;              // the tile kernel waits for this event (counter reset)
};

// Dense arrays + tile search + (flattened, file-index) forest into d_out_label (device memory, n x u32).
// *d_counters -> device: [0] this GPU's comparison count.
int hamming_group_enqueue(rh_ctx *ctx, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                          const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                          const HammingPlan &plan, uint32_t *d_out_label, const unsigned long long **d_counters);
// world x n forests (device memory) -> canonical labels (device memory)
int uf_merge_enqueue(rh_ctx *ctx, const uint32_t *d_parents, int world, int64_t n, uint32_t *d_out_label);
// reads the tile kernel's event pair into ctx->last_ms once the stream has been synchronised
int finish_timing(rh_ctx *ctx);

}  // namespace rh
