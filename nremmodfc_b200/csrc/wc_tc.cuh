// tcgen05 variant of the batched integrator (placeholder until the kernel lands).
#pragma once
#include "wc_batch.cuh"

namespace nrem {
static int launch_wc_tc(int kernel, const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    (void)kernel; (void)A; (void)tiles; (void)st;
    return fail(NREM_ERR_UNSUPPORTED, "tcgen05 integrator not built%s%s");
}
}  // namespace nrem
