// minhash.cu - L4 similarity hashing: MinHash signatures (spec: minhash_compute,
// README.md:2578-2597; 128 x MurmurHash3_x86_32 over 4-byte shingles at every byte offset,
// running minimum from 0xFFFFFFFF) and LSH band keys (README.md:2231-2235).
//
// INT32-issue bound: (len-3) x n_perm murmur evaluations per chunk.  The seed-independent key
// mixing (k*c1, rotl 15, *c2) is hoisted out of the seed loop and, because
// rotl(seed ^ k, 13) == rotl(seed,13) ^ rotl(k,13), so is the first rotate: per (shingle, seed)
// the inner loop is 1 xor + 3 multiplies + 3 xor-shifts + 1 min.
// One warp owns one chunk at a time (pulled from a global counter); each lane keeps
// n_perm/32 running minima in registers, shingles are read coalesced 32 at a time and
// broadcast with shuffles.
#include "ctx.cuh"

namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

template <int PER_LANE>
__global__ void __launch_bounds__(128)
minhash_kernel(const uint8_t* __restrict__ data, uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t n_chunks,
               const uint32_t* __restrict__ seeds, uint32_t* __restrict__ sig,
               unsigned long long* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31;
    uint32_t rs[PER_LANE];
#pragma unroll
    for (int p = 0; p < PER_LANE; p++) rs[p] = rotl32(seeds[p * 32 + lane], 13);
    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(counter, 1ull);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= n_chunks) break;
        const uint64_t s = j ? cuts[j - 1] : start0;
        const uint64_t e = cuts[j];
        uint32_t mn[PER_LANE];
#pragma unroll
        for (int p = 0; p < PER_LANE; p++) mn[p] = 0xFFFFFFFFu;
        const uint64_t n_sh = e - s >= 4 ? e - s - 3 : 0;
        for (uint64_t b0 = 0; b0 < n_sh; b0 += 32) {
            // lane's shingle: little-endian u32 at byte offset s + b0 + lane (unaligned)
            const uint64_t q = b0 + lane;
            uint32_t rk = 0;
            if (q < n_sh) {
                const uint8_t* p = data + s + q;
                const unsigned k = (unsigned)((uintptr_t)p & 3);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(p - k);
                uint32_t lo = __ldg(wp);
                uint32_t hi = k ? __ldg(wp + 1) : 0u;
                uint32_t key = __funnelshift_r(lo, hi, k * 8);
                key *= 0xcc9e2d51u;
                key = rotl32(key, 15);
                key *= 0x1b873593u;
                rk = rotl32(key, 13);
            }
            const uint64_t left = n_sh - b0;
            const int cnt = left < 32 ? (int)left : 32;
#pragma unroll 4
            for (int i = 0; i < cnt; i++) {
                const uint32_t r = __shfl_sync(0xffffffffu, rk, i);
#pragma unroll
                for (int p = 0; p < PER_LANE; p++) {
                    uint32_t h = (rs[p] ^ r) * 5u + 0xe6546b64u;
                    h = h ^ 4u ^ (h >> 16);
                    h *= 0x85ebca6bu;
                    h ^= h >> 13;
                    h *= 0xc2b2ae35u;
                    h ^= h >> 16;
                    mn[p] = min(mn[p], h);
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

template <int PER_LANE>
int launch_minhash(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, uint64_t n_chunks,
                   const uint32_t* d_seeds, uint32_t* d_sig, unsigned long long* counter, cudaStream_t st) {
    uint64_t blocks = div_up64(n_chunks, 4);
    const uint64_t max_blocks = (uint64_t)ctx->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    HT_BEGIN(ctx, HT_MINHASH, st);
    KL(ctx);
    minhash_kernel<PER_LANE><<<(unsigned)blocks, 128, 0, st>>>(d_data, start0, d_cuts, n_chunks, d_seeds, d_sig,
                                                               counter);
    HMSE_LAUNCH_CHECK(ctx);
    HT_END(ctx, HT_MINHASH, st);
    return HMSE_OK;
}

}  // namespace

HMSE_API int hmse_minhash(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts,
                            uint64_t n_chunks, const uint32_t* d_seeds, uint32_t n_perm, uint32_t* d_sig,
                            void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_perm == 0 || n_perm % 32 || n_perm > 256) HMSE_FAIL(ctx, HMSE_E_INVAL, "n_perm must be a multiple of 32, <= 256");
    if (n_chunks == 0) return HMSE_OK;
    if (!d_data || !d_cuts || !d_seeds || !d_sig) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_minhash: null pointer");
    if ((uintptr_t)d_data & 3) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_data must be 4-byte aligned");
    HMSE_SCRATCH(ctx, counter, unsigned long long*, SLOT_MINHASH_MISC, 64);
    HMSE_CUDA(ctx, cudaMemsetAsync(counter, 0, 8, st));
    switch (n_perm / 32) {
        case 1: return launch_minhash<1>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 2: return launch_minhash<2>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 3: return launch_minhash<3>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 4: return launch_minhash<4>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 5: return launch_minhash<5>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 6: return launch_minhash<6>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        case 7: return launch_minhash<7>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
        default: return launch_minhash<8>(ctx, d_data, start0, d_cuts, n_chunks, d_seeds, d_sig, counter, st);
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
