// topk_umma.cu — A4 on the 5th-generation tensor cores (placeholder until the
// tcgen05 kernel lands: reports "unsupported" so AUTO picks the FFMA kernel).
#include "common.cuh"

bool fwav_topk_umma_supported(int, int, int64_t, int64_t) { return false; }

int fwav_launch_topk_umma(fwav_ctx *ctx, const float *, int64_t, const float *, int64_t, int, int,
                          const uint8_t *, int32_t *, float *, cudaStream_t) {
    return fwav_set_error(ctx, FWAV_ERR_UNSUPPORTED, "tensor-core search kernel not built");
}
