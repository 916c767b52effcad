// phash.cu -- 64-bit DCT pHash (phash.rs:48-83).  Device path pending; the bit-level
// dihedral operations live in ctx.cu.
#include "common.cuh"

extern "C" int rh_phash_batch(rh_ctx *ctx, const uint8_t *, int, int64_t, int, int, size_t, size_t, uint64_t *,
                              uint64_t *) {
    if (!ctx) return RH_EINVAL;
    return rh::fail(ctx, RH_EUNSUPPORTED, "rh_phash_batch: not implemented yet");
}
