// sha256.cu - L3 digest: FIPS 180-4 SHA-256 of every raw chunk (spec: README.md:290, 2518-2561;
// the skeleton's mbedtls_sha256 call at README.md:2543).  Integer-ALU bound, no tensor cores.
//
// One hash is strictly serial, so parallelism is across chunks: every LANE owns one chunk at a
// time and pulls the next one from a global counter when it finishes (persistent lanes), which
// keeps warps converged on the 64-round block function while chunk lengths vary 2-32 KiB.
// Full blocks are read as aligned 32-bit words and byte-permuted into big-endian order with one
// PRMT each (the chunk start is byte-unaligned); only the last one or two blocks of a chunk take
// the byte-wise padding path.
#include <stdlib.h>

#include "ctx.cuh"

namespace {

__constant__ uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int r) { return __funnelshift_r(x, x, r); }

// a + b as a * one + b with `one` a kernel argument (always 1): ptxas cannot fold the multiply, so the addition issues
// as an IMAD on the fma pipe.  The block function is otherwise all alu-pipe work (21.5 alu vs 1.9 fma instructions per
// round, ncu: alu pipe 84 % busy, fma 4 %), and both pipes issue one warp instruction per two cycles.
__device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}

template <int FMA_ADDS>
__device__ __forceinline__ void sha256_block(uint32_t (&H)[8], uint32_t (&W)[16], uint32_t one) {
    uint32_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = W[(i + 1) & 15], w2 = W[(i + 14) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            if (FMA_ADDS >= 1) W[i & 15] = madd(madd(W[i & 15], s0, one), madd(W[(i + 9) & 15], s1, one), one);
            else W[i & 15] = W[i & 15] + s0 + W[(i + 9) & 15] + s1;
        }
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t1, t2, en, an;
        if (FMA_ADDS >= 2) {
            const uint32_t x = madd(h, W[i & 15] + K256[i], one);   // off the critical path: h and W are old values
            t1 = madd(madd(S1, ch, one), x, one);
            t2 = madd(S0, mj, one);
            en = madd(d, t1, one);
            an = madd(t1, t2, one);
        } else {
            t1 = h + S1 + ch + K256[i] + W[i & 15];
            t2 = S0 + mj;
            en = d + t1;
            an = t1 + t2;
        }
        h = g;
        g = f;
        f = e;
        e = en;
        d = c;
        c = b;
        b = a;
        a = an;
    }
    H[0] += a; H[1] += b; H[2] += c; H[3] += d; H[4] += e; H[5] += f; H[6] += g; H[7] += h;
}

constexpr int SHA_THREADS = 128;

template <int FMA_ADDS>
__global__ void __launch_bounds__(SHA_THREADS)
sha256_kernel(const uint8_t* __restrict__ data, uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t n_chunks,
              uint8_t* __restrict__ digests, unsigned long long* __restrict__ counter, uint32_t one,
              const uint32_t* __restrict__ order) {
    uint32_t H[8], W[16];
    uint64_t j = 0, s = 0, len = 0, blk = 0, nblk = 0;
    bool have = false, done = false;
    for (;;) {
        if (!have && !done) {
            j = atomicAdd(counter, 1ull);
            if (j >= n_chunks) {
                done = true;
            } else {
                if (order) j = order[j];   // longest chunks first when the lanes get only a few chunks each
                s = j ? cuts[j - 1] : start0;
                len = cuts[j] - s;
                nblk = (len + 9 + 63) >> 6;
                blk = 0;
                H[0] = 0x6a09e667; H[1] = 0xbb67ae85; H[2] = 0x3c6ef372; H[3] = 0xa54ff53a;
                H[4] = 0x510e527f; H[5] = 0x9b05688c; H[6] = 0x1f83d9ab; H[7] = 0x5be0cd19;
                have = true;
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (have) {
            const uint64_t off = blk << 6;
            if (off + 68 <= len) {
                // full block with >= 4 bytes of the same chunk after it: aligned words + PRMT
                const uint8_t* p = data + s + off;
                const unsigned k = (unsigned)((uintptr_t)p & 3);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(p - k);
                const uint32_t sel = 0x0123u + 0x1111u * k;
                uint32_t lo = __ldg(wp);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t hi = __ldg(wp + i + 1);
                    W[i] = __byte_perm(lo, hi, sel);
                    lo = hi;
                }
            } else {
                // tail: message bytes, 0x80, zeros, 64-bit big-endian bit length in the last block
                const uint8_t* p = data + s;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t w = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        uint64_t o = off + 4 * i + b;
                        uint32_t v = o < len ? (uint32_t)p[o] : (o == len ? 0x80u : 0u);
                        w = (w << 8) | v;
                    }
                    W[i] = w;
                }
                if (blk == nblk - 1) {
                    const uint64_t bits = len << 3;
                    W[14] = (uint32_t)(bits >> 32);
                    W[15] = (uint32_t)bits;
                }
            }
            sha256_block<FMA_ADDS>(H, W, one);
            blk++;
            if (blk == nblk) {
                uint4* o = reinterpret_cast<uint4*>(digests + (j << 5));
                o[0] = make_uint4(__byte_perm(H[0], 0, 0x0123), __byte_perm(H[1], 0, 0x0123),
                                  __byte_perm(H[2], 0, 0x0123), __byte_perm(H[3], 0, 0x0123));
                o[1] = make_uint4(__byte_perm(H[4], 0, 0x0123), __byte_perm(H[5], 0, 0x0123),
                                  __byte_perm(H[6], 0, 0x0123), __byte_perm(H[7], 0, 0x0123));
                have = false;
            }
        }
    }
}

// Longest-first work order for batches that give every lane only a few chunks (a piece of a stream): chunk indices
// grouped by length class (2 KiB steps), longest class first, so the lanes that finish last hold short chunks.  The
// order inside a class is whatever the atomics give - the digests do not depend on it.
constexpr int LEN_CLASSES = 33;
__device__ __forceinline__ uint32_t len_class(uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t j) {
    const uint64_t len = cuts[j] - (j ? cuts[j - 1] : start0);
    return len >> 11 < LEN_CLASSES - 1 ? (uint32_t)(len >> 11) : LEN_CLASSES - 1;
}
// (one atomic per class and warp: the lanes of a class are counted with a match)
__global__ void sha_class_hist_kernel(uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t n, uint32_t* __restrict__ hist) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = j < n ? len_class(start0, cuts, j) : 0xFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, c);
    if (j < n && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[c], (uint32_t)__popc(peers));
}
__global__ void sha_class_scan_kernel(uint32_t* __restrict__ hist) {   // cursor[c] = chunks in longer classes
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = LEN_CLASSES - 1; c >= 0; c--) {
            const uint32_t k = hist[c];
            hist[c] = run;
            run += k;
        }
    }
}
__global__ void sha_class_scatter_kernel(uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t n, uint32_t* __restrict__ cursor,
                                         uint32_t* __restrict__ order) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t c = j < n ? len_class(start0, cuts, j) : 0xFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (j < n && lane == (uint32_t)leader) base = atomicAdd(&cursor[c], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (j < n) order[base + __popc(peers & ((1u << lane) - 1))] = (uint32_t)j;
}

}  // namespace

HMSE_API int hmse_digest(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts,
                           uint64_t n_chunks, uint8_t* d_digests, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_chunks == 0) return HMSE_OK;
    if (!d_data || !d_cuts || !d_digests) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_digest: null pointer");
    if ((uintptr_t)d_digests & 15) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_digests must be 16-byte aligned");
    if ((uintptr_t)d_data & 3) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_data must be 4-byte aligned");
    if (n_chunks > 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_digest: n_chunks exceeds 2^32");
    // persistent lanes: enough warps to fill the machine, never more lanes than chunks
    uint64_t blocks = div_up64(n_chunks, SHA_THREADS);
    uint64_t max_blocks = (uint64_t)ctx->sm_count * 8;
    // A batch that gives the lanes of a full machine fewer than ~3 chunks each (a piece of a stream: 2 GiB = 229 k chunks
    // for 151 k lanes) ends when its longest chunks end, and a lane runs slower the more warps share its scheduler (the
    // alu pipe issues one warp instruction per two cycles: 8 warps per scheduler = one round instruction per 16 cycles).
    // Fewer, faster lanes with ~3 chunks each, longest first, finish such a batch sooner: the alu pipe stays saturated down
    // to about two warps per scheduler.
    {
        static int per_sm_override = -1;   // HMSE_SHA_BLOCKS_PER_SM: experiments (tools/sha_pieces.py)
        if (per_sm_override < 0) {
            const char* e = getenv("HMSE_SHA_BLOCKS_PER_SM");
            per_sm_override = e ? atoi(e) : 0;
        }
        if (per_sm_override > 0) max_blocks = (uint64_t)ctx->sm_count * (uint64_t)per_sm_override;
        else {
            uint64_t per_sm = div_up64(div_up64(n_chunks, 3), (uint64_t)ctx->sm_count * SHA_THREADS);
            if (per_sm < 2) per_sm = 2;
            if (per_sm < 8) max_blocks = (uint64_t)ctx->sm_count * per_sm;
        }
    }
    if (blocks > max_blocks) blocks = max_blocks;
    const uint64_t lanes = max_blocks * SHA_THREADS;
    // up to six chunks per lane the stream order leaves long chunks for the end: order them longest first
    const bool lpt = n_chunks > lanes && n_chunks < 6 * lanes;
    // misc: [counter u64][class counters u32 x 64][order u32 n]
    HMSE_SCRATCH(ctx, counter, unsigned long long*, SLOT_SHA_MISC, 8 + 256 + (lpt ? n_chunks * 4 : 0) + 64);
    uint32_t* hist = reinterpret_cast<uint32_t*>(counter + 1);
    uint32_t* order = lpt ? hist + 64 : nullptr;
    HMSE_CUDA(ctx, cudaMemsetAsync(counter, 0, 8 + 256, st));
    HT_BEGIN(ctx, HT_SHA, st);
    if (lpt) {
        KL(ctx);
        sha_class_hist_kernel<<<(unsigned)div_up64(n_chunks, 256), 256, 0, st>>>(start0, d_cuts, n_chunks, hist);
        KL(ctx);
        sha_class_scan_kernel<<<1, 32, 0, st>>>(hist);
        KL(ctx);
        sha_class_scatter_kernel<<<(unsigned)div_up64(n_chunks, 256), 256, 0, st>>>(start0, d_cuts, n_chunks, hist, order);
    }
    KL(ctx);
    // FMA_ADDS = 2 (message schedule and round additions as IMADs) measured on B200: 14.6 -> 13.1 ms per 10 GB
    // (687 -> 763 GB/s); moving the W + K addition as well (a second IMAD) was slower again (13.5 ms).
    sha256_kernel<2><<<(unsigned)blocks, SHA_THREADS, 0, st>>>(d_data, start0, d_cuts, n_chunks, d_digests, counter, 1u, order);
    HMSE_LAUNCH_CHECK(ctx);
    HT_END(ctx, HT_SHA, st);
    return HMSE_OK;
}
