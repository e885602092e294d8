// comm.cu - the multi-GPU exchange steps of the hot path behind the C ABI (north_star: "the
// corpus shards by byte range across the GPUs, with boundary resync at shard edges ... global dedup and LSH bucketing
// use a GPU hash table partitioned by digest / band prefix with NCCL all-to-all over NVLink").  One process per GPU.
//
//   hmse_chunk_sharded   scan + speculative resolve of the local byte-range shard, then the exits travel with one
//                        ncclAllGather per round and every rank re-resolves incrementally from its true entry until no
//                        entry changes (normally one round) - the cut list of the single stream, shard by shard.
//   hmse_dedup_global    partition {digest, gid} records by owner = le32(digest) % world, ncclSend/ncclRecv all-to-all,
//                        owner table (smallest gid wins), all-to-all back, scatter: canon / is_first over the WHOLE stream.
//   hmse_lsh_exchange    band b belongs to rank b % world: every rank sends each owner the columns of its key matrix that
//                        the owner holds; rows arrive in global id order, ready for hmse_lsh_buckets.
//
// Everything is queued on the caller's stream; the host learns counts through ONE mapped-mailbox read per call (the
// Python layer used three .item() round trips per stitch round and two more per exchange).  NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy the process already holds - torch's - or the system's), so the library itself loads
// on machines without NCCL and single-GPU callers never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "ctx.cuh"

namespace {

struct NcclApi {
    void* handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*GetVersion)(int*);
};
NcclApi g_nccl;
int g_nccl_state = 0;  // 0 not tried, 1 loaded, -1 failed
char g_nccl_err[256];

const NcclApi* nccl_api() {
    if (g_nccl_state) return g_nccl_state > 0 ? &g_nccl : nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        snprintf(g_nccl_err, sizeof(g_nccl_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
        g_nccl_state = -1;
        return nullptr;
    }
    g_nccl.handle = h;
    bool ok = true;
#define NCCL_SYM(field, name)                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                       \
    if (!g_nccl.field) {                                             \
        snprintf(g_nccl_err, sizeof(g_nccl_err), "libnccl lacks %s", name); \
        ok = false;                                                  \
    }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(CommCount, "ncclCommCount")
    NCCL_SYM(CommUserRank, "ncclCommUserRank")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
    NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetVersion, "ncclGetVersion")
#undef NCCL_SYM
    g_nccl_state = ok ? 1 : -1;
    return ok ? &g_nccl : nullptr;
}

#define HMSE_NCCL(ctx, api, call)                                                                       \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,         \
                     (api)->GetErrorString(r__));                                                       \
            return HMSE_E_NCCL;                                                                         \
        }                                                                                               \
    } while (0)

struct Comm {
    const NcclApi* api;
    ncclComm_t comm;
    int world, rank;
};

// Resolves the communicator of a call: `comm` (the caller's ncclComm_t) or, when null, the one hmse_comm_init made.
int get_comm(hmse_ctx* ctx, void* comm, Comm* out) {
    const NcclApi* api = nccl_api();
    if (!api) HMSE_FAIL(ctx, HMSE_E_NCCL, "NCCL is not available: %s", g_nccl_err);
    ncclComm_t c = comm ? (ncclComm_t)comm : (ncclComm_t)ctx->comm;
    if (!c) HMSE_FAIL(ctx, HMSE_E_INVAL, "no communicator: pass an ncclComm_t or call hmse_comm_init first");
    out->api = api;
    out->comm = c;
    HMSE_NCCL(ctx, api, api->CommCount(c, &out->world));
    HMSE_NCCL(ctx, api, api->CommUserRank(c, &out->rank));
    if (out->world < 1 || out->world > HMSE_MAX_WORLD)
        HMSE_FAIL(ctx, HMSE_E_INVAL, "communicators of 1..%d ranks are supported (got %d)", HMSE_MAX_WORLD, out->world);
    return HMSE_OK;
}

__global__ void put_u64_kernel(uint64_t* __restrict__ dst, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    if (threadIdx.x == 0) {
        dst[0] = a;
        dst[1] = b;
        dst[2] = c;
        dst[3] = d;
    }
}

// dst[r][c] = keys[r][first + c * step]  for c < cols (the columns one owner holds), row-major
__global__ void lsh_pack_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t bands, uint32_t first, uint32_t step,
                                uint32_t cols, uint64_t* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const uint64_t r = i / cols;
    const uint32_t c = (uint32_t)(i - r * cols);
    dst[i] = keys[r * bands + first + c * step];
}

// Four u64 per rank all-gathered into the mailbox: out[r*4 + q] on the host after the call (synchronises `stream`).
int allgather4(hmse_ctx* ctx, const Comm& cm, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t* out, cudaStream_t st) {
    HMSE_SCRATCH(ctx, buf, uint64_t*, SLOT_COMM_SMALL, (size_t)(cm.world + 1) * 4 * 8);
    KL(ctx);
    put_u64_kernel<<<1, 32, 0, st>>>(buf, a, b, c, d);
    HMSE_LAUNCH_CHECK(ctx);
    HMSE_NCCL(ctx, cm.api, cm.api->AllGather(buf, buf + 4, 4, ncclUint64, cm.comm, st));
    if (int rc = hmse_mail(ctx, 0, buf + 4, (uint32_t)cm.world * 8, st)) return rc;
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < cm.world * 4; i++) out[i] = ctx->pinned[i];
    return HMSE_OK;
}

}  // namespace

HMSE_API int hmse_comm_unique_id(uint8_t* out128) {
    const NcclApi* api = nccl_api();
    if (!api || !out128) return api ? HMSE_E_INVAL : HMSE_E_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return HMSE_E_NCCL;
    static_assert(sizeof(id) == HMSE_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(out128, &id, sizeof(id));
    return HMSE_OK;
}

HMSE_API int hmse_comm_init(hmse_ctx* ctx, const uint8_t* id128, int world, int rank) {
    if (!ctx) return HMSE_E_INVAL;
    const NcclApi* api = nccl_api();
    if (!api) HMSE_FAIL(ctx, HMSE_E_NCCL, "NCCL is not available: %s", g_nccl_err);
    if (!id128 || world < 1 || world > HMSE_MAX_WORLD || rank < 0 || rank >= world)
        HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_comm_init: bad id / world / rank");
    if (ctx->comm) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_comm_init: this ctx already has a communicator");
    HMSE_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    HMSE_NCCL(ctx, api, api->CommInitRank(&c, world, id, rank));
    ctx->comm = c;
    ctx->comm_owned = 1;
    return HMSE_OK;
}

HMSE_API int hmse_comm_destroy(hmse_ctx* ctx) {
    if (!ctx) return HMSE_E_INVAL;
    if (ctx->comm && ctx->comm_owned) {
        const NcclApi* api = nccl_api();
        if (api) api->CommDestroy((ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_owned = 0;
    return HMSE_OK;
}

HMSE_API int hmse_comm_info(hmse_ctx* ctx, void* comm, int* world, int* rank, int* nccl_version) {
    if (!ctx) return HMSE_E_INVAL;
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    if (world) *world = cm.world;
    if (rank) *rank = cm.rank;
    if (nccl_version) cm.api->GetVersion(nccl_version);
    return HMSE_OK;
}

HMSE_API int hmse_allgather_u64(hmse_ctx* ctx, void* comm, const uint64_t* vals4, uint64_t* out, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (!vals4 || !out) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_allgather_u64: null pointer");
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    return allgather4(ctx, cm, vals4[0], vals4[1], vals4[2], vals4[3], out, (cudaStream_t)stream);
}

HMSE_API int hmse_chunk_sharded(hmse_ctx* ctx, void* comm, const uint8_t* d_data, uint64_t n_own, uint64_t n_avail, int eof,
                                const hmse_cdc_cfg* cfg, uint64_t* d_cuts, uint64_t cap, uint64_t* n_cuts, uint64_t* entry,
                                uint64_t* id_base, uint64_t* n_total, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (!n_cuts || !entry) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_chunk_sharded: n_cuts / entry is null");
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = hmse_chunk_scan(ctx, d_data, n_avail, cfg, stream)) return rc;
    const uint64_t own = eof ? n_avail : n_own;
    uint64_t cur_entry = 0, exit_off = 0;
    if (int rc = hmse_chunk_resolve(ctx, d_data, n_own, n_avail, eof, 0, d_cuts, cap, n_cuts, &exit_off, stream)) return rc;
    uint64_t all[4 * HMSE_MAX_WORLD], entries[HMSE_MAX_WORLD] = {0};   // entries[r]: the entry rank r resolved from last
    int rounds = 0;
    HT_BEGIN(ctx, HT_EXCHANGE, st);
    for (;;) {
        // one collective per round: (exit, owned bytes, -, chunk count) of every rank.  The entry of rank r is the exit of
        // rank r-1 minus its owned bytes; every rank derives ALL entries from the same gathered values, so all ranks agree
        // on whether anybody has to re-resolve - no second collective for the termination test.
        rounds++;
        if (rounds > 64 + cm.world) HMSE_FAIL(ctx, HMSE_E_INVAL, "shard boundary resync did not converge in %d rounds", rounds);
        if (int rc = allgather4(ctx, cm, exit_off, own, 0, *n_cuts, all, st)) return rc;
        bool any = false;
        for (int r = 1; r < cm.world; r++) {
            const uint64_t ex = all[4 * (r - 1)], ow = all[4 * (r - 1) + 1];
            const uint64_t e = ex > ow ? ex - ow : 0;
            if (e != entries[r]) any = true;
            entries[r] = e;
        }
        if (!any) break;
        if (entries[cm.rank] != cur_entry) {
            cur_entry = entries[cm.rank];
            if (int rc = hmse_chunk_resolve(ctx, d_data, n_own, n_avail, eof, cur_entry, d_cuts, cap, n_cuts, &exit_off, stream))
                return rc;
        }
    }
    HT_END(ctx, HT_EXCHANGE, st);
    ctx->comm_rounds = rounds;
    *entry = cur_entry;
    uint64_t base = 0, tot = 0;
    for (int r = 0; r < cm.world; r++) {
        if (r < cm.rank) base += all[4 * r + 3];
        tot += all[4 * r + 3];
    }
    if (id_base) *id_base = base;
    if (n_total) *n_total = tot;
    return HMSE_OK;
}

HMSE_API int hmse_dedup_global(hmse_ctx* ctx, void* comm, const uint8_t* d_digests, uint64_t n, uint64_t id_base,
                               int64_t* d_canon, uint8_t* d_is_first, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (n && (!d_digests || !d_canon)) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_dedup_global: null pointer");
    if (n > 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_dedup_global: n exceeds 2^32");
    const int W = cm.world;
    // send side: records grouped by owner (counts stay on the device: they travel with the all-gather below)
    HMSE_SCRATCH(ctx, send, uint8_t*, SLOT_COMM_SEND, (n + 1) * 40 + (n + 1) * 4);
    uint32_t* perm = reinterpret_cast<uint32_t*>(send + (n + 1) * 40);
    HMSE_SCRATCH(ctx, small, uint64_t*, SLOT_COMM_SMALL, (size_t)(W + 1) * (size_t)W * 8 + 64);
    uint64_t* my_counts = small;          // [W]
    uint64_t* matrix = small + W;         // [W][W]: matrix[r][o] = records rank r sends to owner o
    if (int rc = hmse_dedup_partition_dev(ctx, d_digests, n, id_base, (uint32_t)W, send, perm, my_counts, st)) return rc;
    HT_BEGIN(ctx, HT_EXCHANGE, st);
    HMSE_NCCL(ctx, cm.api, cm.api->AllGather(my_counts, matrix, (size_t)W, ncclUint64, cm.comm, st));
    if ((size_t)W * W * 8 > HMSE_MAILBOX_BYTES) HMSE_FAIL(ctx, HMSE_E_INVAL, "world too large for the mailbox");
    if (int rc = hmse_mail(ctx, 0, matrix, (uint32_t)(W * W * 2), st)) return rc;
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));   // the ONE host round trip of the exchange
    uint64_t scnt[HMSE_MAX_WORLD], rcnt[HMSE_MAX_WORLD], soff[HMSE_MAX_WORLD + 1], roff[HMSE_MAX_WORLD + 1];
    soff[0] = roff[0] = 0;
    for (int p = 0; p < W; p++) {
        scnt[p] = ctx->pinned[(size_t)cm.rank * W + p];
        rcnt[p] = ctx->pinned[(size_t)p * W + cm.rank];
        soff[p + 1] = soff[p] + scnt[p];
        roff[p + 1] = roff[p] + rcnt[p];
    }
    if (soff[W] != n) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_dedup_global: partition counts %llu != n %llu",
                                (unsigned long long)soff[W], (unsigned long long)n);
    const uint64_t m = roff[W];
    HMSE_SCRATCH(ctx, recv, uint8_t*, SLOT_COMM_RECV, (m + 1) * 40 + (m + 1) * 8 + (n + 1) * 8);
    uint64_t* answers = reinterpret_cast<uint64_t*>(recv + (m + 1) * 40);
    uint64_t* reply = answers + (m + 1);
    HMSE_NCCL(ctx, cm.api, cm.api->GroupStart());
    for (int p = 0; p < W; p++) {
        if (scnt[p]) HMSE_NCCL(ctx, cm.api, cm.api->Send(send + soff[p] * 40, scnt[p] * 40, ncclUint8, p, cm.comm, st));
        if (rcnt[p]) HMSE_NCCL(ctx, cm.api, cm.api->Recv(recv + roff[p] * 40, rcnt[p] * 40, ncclUint8, p, cm.comm, st));
    }
    HMSE_NCCL(ctx, cm.api, cm.api->GroupEnd());
    if (int rc = hmse_dedup_records(ctx, recv, m, answers, stream)) return rc;
    HMSE_NCCL(ctx, cm.api, cm.api->GroupStart());
    for (int p = 0; p < W; p++) {
        if (rcnt[p]) HMSE_NCCL(ctx, cm.api, cm.api->Send(answers + roff[p], rcnt[p], ncclUint64, p, cm.comm, st));
        if (scnt[p]) HMSE_NCCL(ctx, cm.api, cm.api->Recv(reply + soff[p], scnt[p], ncclUint64, p, cm.comm, st));
    }
    HMSE_NCCL(ctx, cm.api, cm.api->GroupEnd());
    HT_END(ctx, HT_EXCHANGE, st);
    ctx->comm_stat[0] = (n - scnt[cm.rank]) * 40 + (m - rcnt[cm.rank]) * 8;   // bytes sent to other ranks
    ctx->comm_stat[1] = (m - rcnt[cm.rank]) * 40 + (n - scnt[cm.rank]) * 8;   // bytes received from other ranks
    ctx->comm_stat[2] = m;                                                    // records this rank owns
    ctx->comm_stat[3] = n;
    return hmse_dedup_scatter(ctx, reply, perm, n, id_base, d_canon, d_is_first, stream);
}

HMSE_API int hmse_lsh_exchange(hmse_ctx* ctx, void* comm, const uint64_t* d_keys, uint64_t n, uint32_t bands,
                               uint64_t* d_owned, uint64_t owned_cap_rows, uint64_t* n_total, uint64_t* id_base,
                               uint32_t* bands_owned, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (!n_total) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_exchange: n_total is null");
    if (bands == 0 || bands > 256) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_exchange: bands must be 1..256");
    if (n && !d_keys) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_exchange: null keys");
    const int W = cm.world;
    uint64_t all[4 * HMSE_MAX_WORLD];
    HT_BEGIN(ctx, HT_EXCHANGE, st);
    if (int rc = allgather4(ctx, cm, n, bands, 0, 0, all, st)) return rc;
    uint64_t tot = 0, base = 0;
    for (int r = 0; r < W; r++) {
        if (all[4 * r + 1] != bands) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_exchange: rank %d uses %llu bands, this rank %u", r,
                                               (unsigned long long)all[4 * r + 1], bands);
        if (r < cm.rank) base += all[4 * r];
        tot += all[4 * r];
    }
    const uint32_t mine = (uint32_t)cm.rank < bands ? (bands - (uint32_t)cm.rank + (uint32_t)W - 1) / (uint32_t)W : 0;  // b % W == rank
    *n_total = tot;
    if (id_base) *id_base = base;
    if (bands_owned) *bands_owned = mine;
    if (!d_owned || owned_cap_rows < tot) {
        HT_END(ctx, HT_EXCHANGE, st);
        if (!d_owned && owned_cap_rows == 0) return HMSE_OK;   // size query
        HMSE_FAIL(ctx, HMSE_E_CAPACITY, "hmse_lsh_exchange: d_owned holds %llu rows, %llu needed", (unsigned long long)owned_cap_rows,
                  (unsigned long long)tot);
    }
    // pack: for owner o the columns o, o + W, ... of the local key matrix, row-major
    HMSE_SCRATCH(ctx, send, uint64_t*, SLOT_COMM_SEND, (n * bands + 1) * 8);
    uint64_t soff[HMSE_MAX_WORLD + 1];
    soff[0] = 0;
    for (int o = 0; o < W; o++) {
        const uint32_t cols = (uint32_t)o < bands ? (bands - (uint32_t)o + (uint32_t)W - 1) / (uint32_t)W : 0;
        soff[o + 1] = soff[o] + n * cols;
        if (n * cols) {
            KL(ctx);
            lsh_pack_kernel<<<(unsigned)div_up64(n * cols, 256), 256, 0, st>>>(d_keys, n, bands, (uint32_t)o, (uint32_t)W, cols, send + soff[o]);
            HMSE_LAUNCH_CHECK(ctx);
        }
    }
    HMSE_NCCL(ctx, cm.api, cm.api->GroupStart());
    uint64_t roff = 0, sent = 0, got = 0;
    for (int p = 0; p < W; p++) {
        const uint64_t sc = soff[p + 1] - soff[p], rc_ = all[4 * p] * mine;
        if (sc) HMSE_NCCL(ctx, cm.api, cm.api->Send(send + soff[p], sc, ncclUint64, p, cm.comm, st));
        if (rc_) HMSE_NCCL(ctx, cm.api, cm.api->Recv(d_owned + roff, rc_, ncclUint64, p, cm.comm, st));
        roff += rc_;
        if (p != cm.rank) {
            sent += sc * 8;
            got += rc_ * 8;
        }
    }
    HMSE_NCCL(ctx, cm.api, cm.api->GroupEnd());
    HT_END(ctx, HT_EXCHANGE, st);
    ctx->comm_stat[0] = sent;
    ctx->comm_stat[1] = got;
    ctx->comm_stat[2] = tot;
    ctx->comm_stat[3] = n;
    return HMSE_OK;
}

HMSE_API int hmse_alltoallv(hmse_ctx* ctx, void* comm, const void* d_send, const uint64_t* send_counts, void* d_recv,
                            uint64_t* recv_counts, uint32_t elem_bytes, uint64_t recv_cap, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    Comm cm;
    if (int rc = get_comm(ctx, comm, &cm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (!send_counts || !recv_counts || elem_bytes == 0) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_alltoallv: null counts / zero element size");
    const int W = cm.world;
    // counts first: matrix[r][o] = elements rank r sends to rank o (one all-gather, one mailbox read)
    HMSE_SCRATCH(ctx, small, uint64_t*, SLOT_COMM_SMALL, (size_t)(W + 1) * (size_t)W * 8 + 64);
    for (int p = 0; p < W; p++) ctx->pinned[1024 + p] = send_counts[p];   // staged through the mapped mailbox (upper half)
    HMSE_CUDA(ctx, cudaMemcpyAsync(small, ctx->pinned_dev + 1024, (size_t)W * 8, cudaMemcpyDefault, st));
    HT_BEGIN(ctx, HT_EXCHANGE, st);
    HMSE_NCCL(ctx, cm.api, cm.api->AllGather(small, small + W, (size_t)W, ncclUint64, cm.comm, st));
    if (int rc = hmse_mail(ctx, 0, small + W, (uint32_t)(W * W * 2), st)) return rc;
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    uint64_t soff = 0, roff = 0, total_r = 0, total_s = 0;
    for (int p = 0; p < W; p++) {
        recv_counts[p] = ctx->pinned[(size_t)p * W + cm.rank];
        total_r += recv_counts[p];
        total_s += send_counts[p];
    }
    // The size query (d_recv == NULL, recv_cap == 0) is a collective of its own and ends here on EVERY rank, whatever the
    // counts: a rank that happened to receive nothing must not go on to send while its peers return (it would wait for
    // receives that are never posted).
    if (!d_recv && recv_cap == 0) {
        HT_END(ctx, HT_EXCHANGE, st);
        return HMSE_OK;
    }
    if (total_r > recv_cap || (total_s && !d_send)) {
        HT_END(ctx, HT_EXCHANGE, st);
        HMSE_FAIL(ctx, HMSE_E_CAPACITY, "hmse_alltoallv: d_recv holds %llu elements, %llu arrive (every rank must size its buffer "
                                        "from the query call: a rank that fails here leaves its peers waiting)",
                  (unsigned long long)recv_cap, (unsigned long long)total_r);
    }
    uint64_t sent = 0, got = 0;
    HMSE_NCCL(ctx, cm.api, cm.api->GroupStart());
    for (int p = 0; p < W; p++) {
        const uint64_t sb = send_counts[p] * elem_bytes, rb = recv_counts[p] * elem_bytes;
        if (sb) HMSE_NCCL(ctx, cm.api, cm.api->Send((const uint8_t*)d_send + soff, sb, ncclUint8, p, cm.comm, st));
        if (rb) HMSE_NCCL(ctx, cm.api, cm.api->Recv((uint8_t*)d_recv + roff, rb, ncclUint8, p, cm.comm, st));
        soff += sb;
        roff += rb;
        if (p != cm.rank) {
            sent += sb;
            got += rb;
        }
    }
    HMSE_NCCL(ctx, cm.api, cm.api->GroupEnd());
    HT_END(ctx, HT_EXCHANGE, st);
    ctx->comm_stat[0] = sent;
    ctx->comm_stat[1] = got;
    ctx->comm_stat[2] = total_r;
    ctx->comm_stat[3] = total_s;
    return HMSE_OK;
}

HMSE_API int hmse_exchange_stats(hmse_ctx* ctx, uint64_t* out4, int* rounds) {
    if (!ctx || !out4) return HMSE_E_INVAL;
    for (int i = 0; i < 4; i++) out4[i] = ctx->comm_stat[i];
    if (rounds) *rounds = ctx->comm_rounds;
    return HMSE_OK;
}
