// pdq_fused.cu -- placeholder until the fused front end lands.
#include "common.cuh"
#include "pdq_tail.cuh"

namespace rh {
int pdq_fused_supported(int, int) { return 0; }
int pdq_fused_run(rh_ctx *ctx, const uint8_t *, int, bool, int64_t, int, int, size_t, size_t, const TailOut &, int64_t,
                  const float *) {
    return fail(ctx, RH_EUNSUPPORTED, "fused PDQ kernel not built");
}
}  // namespace rh
