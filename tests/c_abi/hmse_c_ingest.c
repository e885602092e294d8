/* hmse_c_ingest.c - the C call sequence of INTEGRATION.md as a stand-alone program: a host written in the reference's
 * own language (the spec's skeletons are ESP-IDF C, README.md:2340-2620) drives libhmse_b200.so through include/hmse.h
 * and the CUDA runtime only - no Python, no torch.  TEST INFRASTRUCTURE (tests/test_gpu_c_abi.py builds and runs it and
 * compares every output with the Python binding and the oracle).
 *
 *   hmse_c_ingest <input file> <dictionary file | -> <output file>
 * output: u64 n_chunks, u64 m, u64 blob bytes, then cuts[n] u64, digests[n][32], canon[n] i64, select[m] u64,
 *         offsets[m+1] u64, blob.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmse.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define HK(x) do { int r_ = (x); if (r_ != HMSE_OK) { fprintf(stderr, "%s: %d %s\n", #x, r_, hmse_last_error(ctx)); return 3; } } while (0)

static uint8_t* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* p = (uint8_t*)malloc(*n + 1);
    if (p && fread(p, 1, *n, f) != *n) { free(p); p = NULL; }
    fclose(f);
    return p;
}

/* the Gear table: the splitmix64 stream started at 0x484D5345 (hmse_b200/config.py, oracle/config.py) */
static void fill_gear(uint64_t* g) {
    uint64_t x = 0x484D5345ull;
    for (int i = 0; i < 256; i++) {
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        g[i] = z ^ (z >> 31);
    }
}

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s input dict|- output\n", argv[0]); return 1; }
    size_t n = 0, dn = 0;
    uint8_t* in = slurp(argv[1], &n);
    uint8_t* zd = strcmp(argv[2], "-") ? slurp(argv[2], &dn) : NULL;
    if (!in || (strcmp(argv[2], "-") && !zd)) { fprintf(stderr, "cannot read the inputs\n"); return 1; }

    hmse_ctx* ctx = NULL;
    if (hmse_abi_version() != HMSE_ABI_VERSION || hmse_create(0, &ctx) != HMSE_OK) { fprintf(stderr, "hmse_create failed\n"); return 3; }
    hmse_cdc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.min_size = 2048; cfg.avg_size = 8192; cfg.max_size = 32768;
    cfg.mask_s = 0x0003590703530000ull; cfg.mask_l = 0x0000d90003530000ull;
    fill_gear(cfg.gear);

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const uint64_t cap = n / cfg.min_size + 2;
    uint8_t *d_data, *d_zd = NULL, *d_dig, *d_first, *d_blob;
    uint64_t *d_cuts, *d_sel, *d_offs;
    int64_t* d_canon;
    CK(cudaMalloc((void**)&d_data, n + 64));            /* kernels read whole 16-byte vectors: keep slack behind the data */
    CK(cudaMemsetAsync(d_data, 0, n + 64, st));
    CK(cudaMemcpyAsync(d_data, in, n, cudaMemcpyHostToDevice, st));
    if (dn) { CK(cudaMalloc((void**)&d_zd, dn + 64)); CK(cudaMemcpyAsync(d_zd, zd, dn, cudaMemcpyHostToDevice, st)); }
    CK(cudaMalloc((void**)&d_cuts, cap * 8));
    CK(cudaMalloc((void**)&d_dig, cap * 32));
    CK(cudaMalloc((void**)&d_canon, cap * 8));
    CK(cudaMalloc((void**)&d_first, cap));
    CK(cudaMalloc((void**)&d_sel, cap * 8));
    CK(cudaMalloc((void**)&d_offs, (cap + 1) * 8));

    uint64_t n_cuts = 0, m = 0, total = 0;
    HK(hmse_chunk(ctx, d_data, n, &cfg, d_cuts, cap, &n_cuts, st));                                  /* L2 */
    HK(hmse_digest(ctx, d_data, 0, d_cuts, n_cuts, d_dig, st));                                       /* L3 */
    HK(hmse_dedup(ctx, d_dig, n_cuts, d_canon, d_first, st));
    HK(hmse_dedup_select(ctx, d_first, n_cuts, d_sel, cap, &m, st));
    uint64_t blob_cap = n / 2 + 4096;                                                                /* L1, with the capacity protocol */
    CK(cudaMalloc((void**)&d_blob, blob_cap));
    int rc = hmse_compress(ctx, d_data, 0, d_cuts, d_sel, m, d_zd, (uint32_t)dn, 6, d_blob, blob_cap, d_offs, &total, st);
    if (rc == HMSE_E_CAPACITY) {
        CK(cudaFree(d_blob));
        blob_cap = total;
        CK(cudaMalloc((void**)&d_blob, blob_cap));
        rc = hmse_compress(ctx, d_data, 0, d_cuts, d_sel, m, d_zd, (uint32_t)dn, 6, d_blob, blob_cap, d_offs, &total, st);
    }
    HK(rc);
    CK(cudaStreamSynchronize(st));

    /* error behaviour: bad arguments come back as codes, never as aborts */
    if (hmse_digest(ctx, NULL, 0, d_cuts, n_cuts, d_dig, st) != HMSE_E_INVAL || !strlen(hmse_last_error(ctx))) {
        fprintf(stderr, "null pointer was not rejected\n");
        return 4;
    }

    uint64_t hdr[3] = {n_cuts, m, total};
    uint8_t* out = (uint8_t*)malloc(n_cuts * 56 + (m + 1) * 8 + m * 8 + total + 64);
    size_t o = 0;
    CK(cudaMemcpy(out + o, d_cuts, n_cuts * 8, cudaMemcpyDeviceToHost)); o += n_cuts * 8;
    CK(cudaMemcpy(out + o, d_dig, n_cuts * 32, cudaMemcpyDeviceToHost)); o += n_cuts * 32;
    CK(cudaMemcpy(out + o, d_canon, n_cuts * 8, cudaMemcpyDeviceToHost)); o += n_cuts * 8;
    CK(cudaMemcpy(out + o, d_sel, m * 8, cudaMemcpyDeviceToHost)); o += m * 8;
    CK(cudaMemcpy(out + o, d_offs, (m + 1) * 8, cudaMemcpyDeviceToHost)); o += (m + 1) * 8;
    CK(cudaMemcpy(out + o, d_blob, total, cudaMemcpyDeviceToHost)); o += total;
    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(hdr, 8, 3, f) != 3 || fwrite(out, 1, o, f) != o) { fprintf(stderr, "cannot write the output\n"); return 1; }
    fclose(f);
    printf("chunks %llu unique %llu compressed %llu of %zu bytes, %llu kernels launched\n", (unsigned long long)n_cuts,
           (unsigned long long)m, (unsigned long long)total, n, (unsigned long long)hmse_launch_count(ctx));
    hmse_destroy(ctx);
    return 0;
}
