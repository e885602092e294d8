// ctx.cu - context lifetime and scratch management for libhmse_b200.so.
#include <stdlib.h>

#include <new>

#include "ctx.cuh"

HMSE_API int hmse_abi_version(void) { return HMSE_ABI_VERSION; }

HMSE_API int hmse_create(int device, hmse_ctx** out) {
    if (!out) return HMSE_E_INVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return HMSE_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return HMSE_E_CUDA;
    hmse_ctx* c = new (std::nothrow) hmse_ctx();
    if (!c) return HMSE_E_NOMEM;
    memset(c, 0, sizeof(*c));
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return HMSE_E_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    if (cudaMallocHost((void**)&c->pinned, 4096) != cudaSuccess) {
        delete c;
        return HMSE_E_CUDA;
    }
    *out = c;
    return HMSE_OK;
}

HMSE_API void hmse_destroy(hmse_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < SLOT_COUNT; i++)
        if (ctx->slot[i]) cudaFree(ctx->slot[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    free(ctx->dict_host);
    delete ctx;
}

HMSE_API const char* hmse_last_error(hmse_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

HMSE_API uint64_t hmse_scratch_bytes(hmse_ctx* ctx) {
    uint64_t s = 0;
    if (ctx)
        for (int i = 0; i < SLOT_COUNT; i++) s += ctx->slot_bytes[i];
    return s;
}

void* hmse_scratch(hmse_ctx* ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 256;
    if (ctx->slot_bytes[slot] >= bytes) return ctx->slot[slot];
    if (ctx->slot[slot]) {
        cudaFree(ctx->slot[slot]);
        ctx->slot[slot] = nullptr;
        ctx->slot_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8;  // headroom so slowly growing inputs do not thrash
    want = (want + 255) & ~(size_t)255;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = (bytes + 255) & ~(size_t)255;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        snprintf(ctx->err, sizeof(ctx->err), "scratch slot %d: cudaMalloc(%zu) failed: %s", slot, want,
                 cudaGetErrorString(e));
        return nullptr;
    }
    ctx->slot[slot] = p;
    ctx->slot_bytes[slot] = want;
    return p;
}
