// todo.cu - entry points declared in include/hmse.h whose kernels are not built yet.
#include "ctx.cuh"
HMSE_API int hmse_lsh_buckets(hmse_ctx* ctx, const uint64_t*, uint64_t, uint32_t, uint64_t, uint32_t*, uint64_t*,
                              uint64_t*, void*) {
    HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_buckets: not built yet");
}
HMSE_API uint64_t hmse_compress_bound(uint64_t len) { return len + 5 * (len / 65535 + 1) + 6 + 4 + 16; }
HMSE_API int hmse_compress(hmse_ctx* ctx, const uint8_t*, uint64_t, const uint64_t*, const uint64_t*, uint64_t,
                           const uint8_t*, uint32_t, int, uint8_t*, uint64_t, uint64_t*, uint64_t*, void*) {
    HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_compress: not built yet");
}
HMSE_API int hmse_corpus_lengths(hmse_ctx* ctx, const hmse_corpus_cfg*, const uint32_t*, uint64_t, uint64_t, uint32_t*,
                                 void*) {
    HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_corpus_lengths: not built yet");
}
HMSE_API int hmse_corpus_render(hmse_ctx* ctx, const hmse_corpus_cfg*, const uint8_t*, const uint32_t*, uint64_t,
                                uint64_t, const uint64_t*, uint64_t, uint64_t, uint8_t*, void*) {
    HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_corpus_render: not built yet");
}
