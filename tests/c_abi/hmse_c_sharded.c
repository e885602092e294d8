/* hmse_c_sharded.c - the multi-GPU call sequence of INTEGRATION.md from plain C: one process per GPU (fork before any
 * CUDA call), the ncclUniqueId travels from rank 0 to the others through a pipe, and every rank drives its byte-range
 * shard of ONE input file through hmse_chunk_sharded -> hmse_digest -> hmse_dedup_global -> hmse_dedup_select ->
 * hmse_compress, then MinHash -> hmse_lsh_keys -> hmse_lsh_exchange -> hmse_lsh_buckets.  No Python, no torch, no MPI.
 * TEST INFRASTRUCTURE (tests/test_gpu_c_abi.py builds and runs it on 2 GPUs and compares every output with the oracle
 * run over the whole file).
 *
 *   hmse_c_sharded <world> <input file> <dictionary file | -> <output prefix>
 * rank r writes <prefix>.<r>: u64 { n_chunks, m, blob bytes, entry, id_base, n_total, shard offset, triples, rounds },
 * then cuts[n] u64 (relative to the shard), digests[n][32], canon[n] i64 (global ids), select[m] u64, offsets[m+1] u64,
 * blob, then the LSH triples of the bands this rank owns: band[t] u32, key[t] u64, id[t] u64.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "hmse.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "rank %d: %s: %s\n", rank, #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define HK(x) do { int r_ = (x); if (r_ != HMSE_OK) { fprintf(stderr, "rank %d: %s: %d %s\n", rank, #x, r_, hmse_last_error(ctx)); return 3; } } while (0)

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

static int run_rank(int rank, int world, const uint8_t* id, const char* in_path, const char* dict_path, const char* prefix) {
    size_t total_n = 0, dn = 0;
    uint8_t* in = slurp(in_path, &total_n);
    uint8_t* zd = strcmp(dict_path, "-") ? slurp(dict_path, &dn) : NULL;
    if (!in) { fprintf(stderr, "rank %d: cannot read %s\n", rank, in_path); return 1; }
    hmse_ctx* ctx = NULL;
    if (hmse_create(rank, &ctx) != HMSE_OK) { fprintf(stderr, "rank %d: hmse_create failed\n", rank); return 3; }
    HK(hmse_comm_init(ctx, id, world, rank));
    int w2 = 0, r2 = 0, ver = 0;
    HK(hmse_comm_info(ctx, NULL, &w2, &r2, &ver));
    if (w2 != world || r2 != rank) { fprintf(stderr, "rank %d: communicator reports %d/%d\n", rank, r2, w2); return 4; }

    hmse_cdc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.min_size = 2048; cfg.avg_size = 8192; cfg.max_size = 32768;
    cfg.mask_s = 0x0003590703530000ull; cfg.mask_l = 0x0000d90003530000ull;
    fill_gear(cfg.gear);

    /* contiguous byte-range shards (16-byte aligned starts); a non-final shard carries max_size bytes of look-ahead */
    const uint64_t per = (total_n / (uint64_t)world) & ~(uint64_t)15;
    const uint64_t lo = per * (uint64_t)rank;
    const int eof = rank == world - 1;
    const uint64_t n_own = eof ? total_n - lo : per;
    const uint64_t n_avail = eof ? n_own : n_own + cfg.max_size;

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const uint64_t cap = n_avail / cfg.min_size + 2;
    uint8_t *d_data, *d_zd = NULL, *d_dig, *d_first, *d_blob;
    uint64_t *d_cuts, *d_sel, *d_offs;
    int64_t* d_canon;
    CK(cudaMalloc((void**)&d_data, n_avail + 64));
    CK(cudaMemsetAsync(d_data, 0, n_avail + 64, st));
    CK(cudaMemcpyAsync(d_data, in + lo, n_avail, cudaMemcpyHostToDevice, st));
    if (dn) { CK(cudaMalloc((void**)&d_zd, dn + 64)); CK(cudaMemcpyAsync(d_zd, zd, dn, cudaMemcpyHostToDevice, st)); }
    CK(cudaMalloc((void**)&d_cuts, cap * 8));
    CK(cudaMalloc((void**)&d_dig, cap * 32));
    CK(cudaMalloc((void**)&d_canon, cap * 8));
    CK(cudaMalloc((void**)&d_first, cap));
    CK(cudaMalloc((void**)&d_sel, cap * 8));
    CK(cudaMalloc((void**)&d_offs, (cap + 1) * 8));

    uint64_t n_cuts = 0, entry = 0, id_base = 0, n_total = 0, m = 0, total = 0;
    HK(hmse_chunk_sharded(ctx, NULL, d_data, n_own, n_avail, eof, &cfg, d_cuts, cap, &n_cuts, &entry, &id_base, &n_total, st));
    uint64_t xs[4];
    int rounds = 0;
    HK(hmse_exchange_stats(ctx, xs, &rounds));
    HK(hmse_digest(ctx, d_data, entry, d_cuts, n_cuts, d_dig, st));
    HK(hmse_dedup_global(ctx, NULL, d_dig, n_cuts, id_base, d_canon, d_first, st));
    HK(hmse_dedup_select(ctx, d_first, n_cuts, d_sel, cap, &m, st));
    uint64_t blob_cap = n_avail + 64 * m + 1024;
    CK(cudaMalloc((void**)&d_blob, blob_cap));
    HK(hmse_compress(ctx, d_data, entry, d_cuts, d_sel, m, d_zd, (uint32_t)dn, 6, d_blob, blob_cap, d_offs, &total, st));

    /* similarity: sign the local chunks, exchange the band keys, sort the owned bands over the whole stream */
    uint32_t seeds[128];
    for (int i = 0; i < 128; i++) seeds[i] = (uint32_t)(i + 1);
    uint32_t *d_seeds, *d_sig, *d_band, bands_owned = 0;
    uint64_t *d_keys, *d_owned, *d_key, *d_id, nt2 = 0, ib2 = 0;
    CK(cudaMalloc((void**)&d_seeds, sizeof seeds));
    CK(cudaMemcpyAsync(d_seeds, seeds, sizeof seeds, cudaMemcpyHostToDevice, st));
    CK(cudaMalloc((void**)&d_sig, (n_cuts + 1) * 128 * 4));
    CK(cudaMalloc((void**)&d_keys, (n_cuts + 1) * 32 * 8));
    HK(hmse_minhash(ctx, d_data, entry, d_cuts, n_cuts, d_seeds, 128, d_sig, st));
    HK(hmse_lsh_keys(ctx, d_sig, n_cuts, 32, 4, d_keys, st));
    HK(hmse_lsh_exchange(ctx, NULL, d_keys, n_cuts, 32, NULL, 0, &nt2, &ib2, &bands_owned, st));   /* size query */
    if (nt2 != n_total || ib2 != id_base) { fprintf(stderr, "rank %d: exchange sizes disagree with the chunker\n", rank); return 4; }
    const uint64_t triples = nt2 * bands_owned;
    CK(cudaMalloc((void**)&d_owned, (triples + 1) * 8));
    CK(cudaMalloc((void**)&d_band, (triples + 1) * 4));
    CK(cudaMalloc((void**)&d_key, (triples + 1) * 8));
    CK(cudaMalloc((void**)&d_id, (triples + 1) * 8));
    HK(hmse_lsh_exchange(ctx, NULL, d_keys, n_cuts, 32, d_owned, nt2, &nt2, &ib2, &bands_owned, st));
    if (bands_owned) HK(hmse_lsh_buckets(ctx, d_owned, nt2, bands_owned, 0, d_band, d_key, d_id, st));
    CK(cudaStreamSynchronize(st));

    /* a communicator-less context must refuse, not crash */
    hmse_ctx* lone = NULL;
    if (hmse_create(rank, &lone) != HMSE_OK || hmse_dedup_global(lone, NULL, d_dig, n_cuts, 0, d_canon, d_first, st) != HMSE_E_INVAL) {
        fprintf(stderr, "rank %d: a ctx without a communicator was not rejected\n", rank);
        return 4;
    }
    hmse_destroy(lone);

    char path[4096];
    snprintf(path, sizeof path, "%s.%d", prefix, rank);
    FILE* f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "rank %d: cannot write %s\n", rank, path); return 1; }
    uint64_t hdr[9] = {n_cuts, m, total, entry, id_base, n_total, lo, triples, (uint64_t)rounds};
    fwrite(hdr, 8, 9, f);
    const size_t sizes[9] = {n_cuts * 8, n_cuts * 32, n_cuts * 8, m * 8, (m + 1) * 8, total, triples * 4, triples * 8, triples * 8};
    const void* ptrs[9] = {d_cuts, d_dig, d_canon, d_sel, d_offs, d_blob, d_band, d_key, d_id};
    for (int i = 0; i < 9; i++) {
        if (!sizes[i]) continue;
        void* h = malloc(sizes[i]);
        CK(cudaMemcpy(h, ptrs[i], sizes[i], cudaMemcpyDeviceToHost));
        if (fwrite(h, 1, sizes[i], f) != sizes[i]) { fprintf(stderr, "rank %d: short write\n", rank); return 1; }
        free(h);
    }
    fclose(f);
    printf("rank %d/%d nccl %d: %llu chunks from offset %llu (entry %llu, id base %llu of %llu), %llu stored, %llu compressed bytes, "
           "%llu bucket triples, %d resync rounds\n", rank, world, ver, (unsigned long long)n_cuts, (unsigned long long)lo,
           (unsigned long long)entry, (unsigned long long)id_base, (unsigned long long)n_total, (unsigned long long)m,
           (unsigned long long)total, (unsigned long long)triples, rounds);
    hmse_destroy(ctx);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s world input dict|- output-prefix\n", argv[0]); return 1; }
    const int world = atoi(argv[1]);
    if (world < 1 || world > 8) { fprintf(stderr, "world must be 1..8\n"); return 1; }
    /* fork first (no CUDA state in the parent), then rank 0 draws the id and the parent relays it to the others */
    int up[2], down[8][2];
    if (pipe(up)) return 1;
    pid_t pids[8];
    for (int r = 0; r < world; r++) {
        if (pipe(down[r])) return 1;
        pids[r] = fork();
        if (pids[r] < 0) return 1;
        if (pids[r] == 0) {
            uint8_t id[HMSE_UNIQUE_ID_BYTES];
            if (r == 0) {
                if (hmse_comm_unique_id(id) != HMSE_OK) { fprintf(stderr, "hmse_comm_unique_id failed (no libnccl.so.2?)\n"); _exit(5); }
                if (write(up[1], id, sizeof id) != (ssize_t)sizeof id) _exit(5);
            }
            if (read(down[r][0], id, sizeof id) != (ssize_t)sizeof id) _exit(5);
            const int rc = run_rank(r, world, id, argv[2], argv[3], argv[4]);
            fflush(stdout);
            _exit(rc);
        }
    }
    uint8_t id[HMSE_UNIQUE_ID_BYTES];
    if (read(up[0], id, sizeof id) != (ssize_t)sizeof id) { fprintf(stderr, "no unique id from rank 0\n"); return 5; }
    for (int r = 0; r < world; r++)
        if (write(down[r][1], id, sizeof id) != (ssize_t)sizeof id) return 5;
    int bad = 0;
    for (int r = 0; r < world; r++) {
        int status = 0;
        waitpid(pids[r], &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status)) {
            fprintf(stderr, "rank %d failed (status %d)\n", r, status);
            bad = 1;
        }
    }
    return bad;
}
