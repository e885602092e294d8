// minhash.cu - L4 similarity hashing: MinHash signatures (spec: minhash_compute,
// README.md:2578-2597; 128 x MurmurHash3_x86_32 over 4-byte shingles at every byte offset,
// running minimum from 0xFFFFFFFF) and LSH band keys (README.md:2231-2235).
//
// INT32-issue bound: (len-3) x n_perm murmur evaluations per chunk.  The seed-independent key
// mixing (k*c1, rotl 15, *c2) is hoisted out of the seed loop and, because
// rotl(seed ^ k, 13) == rotl(seed,13) ^ rotl(k,13), so is the first rotate: per (shingle, seed)
// the inner loop is 1 xor + 3 multiplies + 3 xor-shifts + 1 min: 8 alu-pipe instructions (4 LOP3, 3 SHF, 1 VIMNMX)
// and 3 fma-pipe IMADs, so the alu pipe bounds it.  Moving shifts to the fma pipe as IMAD.HI was measured on B200
// and is slower (16.1 -> 15.3 / 14.9 / 12.4 GB/s with one / two / three shifts moved).
// One warp owns one chunk at a time (pulled from a global counter); each lane keeps
// n_perm/32 running minima in registers, shingles are read coalesced 32 at a time and
// broadcast with shuffles, two per step (VIMNMX3 folds both hashes into the running minimum).
// Repeated shingles of a chunk are skipped through a per-warp table of recently seen values.
#include "ctx.cuh"

namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// FB > 0: every warp keeps a direct-mapped table of 2^FB recently seen shingles of the current chunk in shared memory; a
// shingle found there has already been hashed for this chunk and is skipped (the minimum over a set does not care about
// repeats, so signatures are unchanged; a collision only costs a repeated evaluation).  On the wiki corpus 29 % of the
// positions of a chunk repeat an earlier shingle; a 4096-entry table catches two thirds of them.
template <int PER_LANE, int FB>
__global__ void __launch_bounds__(128)
minhash_kernel(const uint8_t* __restrict__ data, uint64_t start0, const uint64_t* __restrict__ cuts, const uint64_t* __restrict__ select,
               uint64_t n_chunks, const uint32_t* __restrict__ seeds, uint32_t* __restrict__ sig,
               unsigned long long* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31;
    extern __shared__ uint32_t s_filter[];
    uint32_t* tab = s_filter + (threadIdx.x >> 5) * (FB ? (1u << FB) : 0u);
    uint32_t rs[PER_LANE];
#pragma unroll
    for (int p = 0; p < PER_LANE; p++) rs[p] = rotl32(seeds[p * 32 + lane], 13);
    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(counter, 1ull);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= n_chunks) break;
        const uint64_t c = select ? select[j] : j;   // signature row j belongs to chunk c
        const uint64_t s = c ? cuts[c - 1] : start0;
        const uint64_t e = cuts[c];
        uint32_t mn[PER_LANE];
#pragma unroll
        for (int p = 0; p < PER_LANE; p++) mn[p] = 0xFFFFFFFFu;
        const uint64_t n_sh = e - s >= 4 ? e - s - 3 : 0;
        if (FB) {   // empty table: slot k holds a value that does not map to slot k (0 maps to slot 0, 1 does not)
            __syncwarp();
            for (uint32_t k = lane; k < (1u << FB); k += 32) tab[k] = k ? 0u : 1u;
            __syncwarp();
        }
        // lane's shingle of a batch: little-endian u32 at byte offset s + b0 + lane (unaligned); the next batch's is
        // fetched before this batch is hashed, so its latency hides behind ~1400 warp instructions
        auto load_key = [&](uint64_t q) -> uint32_t {
            if (q >= n_sh) return 0u;
            const uint8_t* p = data + s + q;
            const unsigned k = (unsigned)((uintptr_t)p & 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(p - k);
            const uint32_t lo = __ldg(wp);
            const uint32_t hi = k ? __ldg(wp + 1) : 0u;
            return __funnelshift_r(lo, hi, k * 8);
        };
        uint32_t key_next = load_key(lane);
        for (uint64_t b0 = 0; b0 < n_sh; b0 += 32) {
            const uint64_t q = b0 + lane;
            uint32_t key = key_next;
            key_next = load_key(q + 32);
            bool fresh = q < n_sh;
            if (FB && fresh) {
                const uint32_t slot = (key * 0x9E3779B1u) >> (32 - (FB ? FB : 1));
                // (a hit means some lane inserted this very value for this chunk and hashes or hashed it: no ordering
                // between the lookups and the insertions of a batch is needed)
                fresh = tab[slot] != key;
                if (fresh) tab[slot] = key;
            }
            key *= 0xcc9e2d51u;
            key = rotl32(key, 15);
            key *= 0x1b873593u;
            const uint32_t rk = rotl32(key, 13);
            // two shingles per step: their hashes and the running minimum meet in one three-input VIMNMX3
            // (an odd tail repeats its last shingle - the minimum is idempotent)
            uint32_t todo = __ballot_sync(0xffffffffu, fresh);
            // the pair of the next step is fetched (two shuffles) while the current pair is hashed
            auto next_pair = [&](uint32_t& ra, uint32_t& rb) {
                const int ia = __ffs(todo) - 1;   // -1 (todo == 0) reads lane 31: unused
                todo &= todo - 1;
                const int ib = todo ? __ffs(todo) - 1 : ia;
                todo &= todo - 1;
                ra = __shfl_sync(0xffffffffu, rk, ia & 31);
                rb = __shfl_sync(0xffffffffu, rk, ib & 31);
            };
            bool more = todo != 0;
            uint32_t na, nb;
            next_pair(na, nb);
            while (more) {
                const uint32_t ra = na, rb = nb;
                more = todo != 0;
                next_pair(na, nb);
#pragma unroll
                for (int p = 0; p < PER_LANE; p++) {
                    uint32_t ha = (rs[p] ^ ra) * 5u + 0xe6546b64u, hb = (rs[p] ^ rb) * 5u + 0xe6546b64u;
                    ha = ha ^ 4u ^ (ha >> 16);
                    hb = hb ^ 4u ^ (hb >> 16);
                    ha *= 0x85ebca6bu;
                    hb *= 0x85ebca6bu;
                    ha ^= ha >> 13;
                    hb ^= hb >> 13;
                    ha *= 0xc2b2ae35u;
                    hb *= 0xc2b2ae35u;
                    ha ^= ha >> 16;
                    hb ^= hb >> 16;
                    mn[p] = __vimin3_u32(mn[p], ha, hb);
                }
            }
        }
        uint32_t* o = sig + j * (uint64_t)(PER_LANE * 32);
#pragma unroll
        for (int p = 0; p < PER_LANE; p++) o[p * 32 + lane] = mn[p];
    }
}

__global__ void lsh_keys_kernel(const uint32_t* __restrict__ sig, uint64_t n, uint32_t bands, uint32_t rows,
                                uint64_t* __restrict__ keys) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * bands) return;
    const uint32_t* p = sig + t * rows;  // chunk-major, band-minor: same linear order as keys
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t r = 0; r < rows; r++) {
        uint32_t v = p[r];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            h = (h ^ ((v >> (8 * b)) & 0xffu)) * 0x100000001B3ull;
        }
    }
    keys[t] = h;
}

template <int PER_LANE, int FB>
int launch_minhash_fb(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, const uint64_t* d_select,
                   uint64_t n_chunks, const uint32_t* d_seeds, uint32_t* d_sig, unsigned long long* counter, cudaStream_t st) {
    uint64_t blocks = div_up64(n_chunks, 4);
    const size_t smem = FB ? (size_t)4 * 4 * (1u << FB) : 0;   // four warps, one table each
    const uint64_t per_sm = FB ? (200u * 1024) / smem : 16;
    const uint64_t max_blocks = (uint64_t)ctx->sm_count * (per_sm < 16 ? per_sm : 16);
    if (blocks > max_blocks) blocks = max_blocks;
    if (smem > 48 * 1024)
        HMSE_CUDA(ctx, cudaFuncSetAttribute(minhash_kernel<PER_LANE, FB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HT_BEGIN(ctx, HT_MINHASH, st);
    KL(ctx);
    minhash_kernel<PER_LANE, FB><<<(unsigned)blocks, 128, smem, st>>>(d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter);
    HMSE_LAUNCH_CHECK(ctx);
    HT_END(ctx, HT_MINHASH, st);
    return HMSE_OK;
}

template <int PER_LANE>
int launch_minhash(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, const uint64_t* d_select,
                   uint64_t n_chunks, const uint32_t* d_seeds, uint32_t* d_sig, unsigned long long* counter, cudaStream_t st) {
    // 4096-entry tables (16 KiB per warp, three CTAs per SM) measured best on B200: 16.3 (no table) / 19.7 (2048) /
    // 20.6 (4096) GB/s; 8192 entries leave one CTA per SM and halve the rate.
    return launch_minhash_fb<PER_LANE, 12>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
}

}  // namespace

static int minhash_impl(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, const uint64_t* d_select,
                        uint64_t n_chunks, const uint32_t* d_seeds, uint32_t n_perm, uint32_t* d_sig, void* stream);

HMSE_API int hmse_minhash(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts,
                            uint64_t n_chunks, const uint32_t* d_seeds, uint32_t n_perm, uint32_t* d_sig,
                            void* stream) {
    return minhash_impl(ctx, d_data, start0, d_cuts, nullptr, n_chunks, d_seeds, n_perm, d_sig, stream);
}

HMSE_API int hmse_minhash_select(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts,
                                 const uint64_t* d_select, uint64_t m, const uint32_t* d_seeds, uint32_t n_perm,
                                 uint32_t* d_sig, void* stream) {
    if (ctx && m && !d_select) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_minhash_select: null d_select");
    return minhash_impl(ctx, d_data, start0, d_cuts, d_select, m, d_seeds, n_perm, d_sig, stream);
}

static int minhash_impl(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, const uint64_t* d_select,
                        uint64_t n_chunks, const uint32_t* d_seeds, uint32_t n_perm, uint32_t* d_sig, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_perm == 0 || n_perm % 32 || n_perm > 256) HMSE_FAIL(ctx, HMSE_E_INVAL, "n_perm must be a multiple of 32, <= 256");
    if (n_chunks == 0) return HMSE_OK;
    if (!d_data || !d_cuts || !d_seeds || !d_sig) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_minhash: null pointer");
    if ((uintptr_t)d_data & 3) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_data must be 4-byte aligned");
    HMSE_SCRATCH(ctx, counter, unsigned long long*, SLOT_MINHASH_MISC, 64);
    HMSE_CUDA(ctx, cudaMemsetAsync(counter, 0, 8, st));
    switch (n_perm / 32) {
        case 1: return launch_minhash<1>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 2: return launch_minhash<2>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 3: return launch_minhash<3>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 4: return launch_minhash<4>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 5: return launch_minhash<5>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 6: return launch_minhash<6>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        case 7: return launch_minhash<7>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
        default: return launch_minhash<8>(ctx, d_data, start0, d_cuts, d_select, n_chunks, d_seeds, d_sig, counter, st);
    }
}

HMSE_API int hmse_lsh_keys(hmse_ctx* ctx, const uint32_t* d_sig, uint64_t n, uint32_t bands, uint32_t rows,
                             uint64_t* d_keys, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (bands == 0 || rows == 0) HMSE_FAIL(ctx, HMSE_E_INVAL, "bands and rows must be positive");
    if (n == 0) return HMSE_OK;
    if (!d_sig || !d_keys) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_keys: null pointer");
    KL(ctx);
    lsh_keys_kernel<<<(unsigned)div_up64(n * bands, 256), 256, 0, (cudaStream_t)stream>>>(d_sig, n, bands, rows,
                                                                                         d_keys);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}
