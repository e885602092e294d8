// delta.cu - L4 delta coding against LSH-selected bases (README.md:1328 "if LSH match -> compute binary
// difference, store delta only if size <= 20 % of original chunk"; flow README.md:1555-1570; Appendix A
// README.md:2160-2198; op-list example README.md:1402-1412).  The spec names xdelta3 / bsdiff without a byte
// format; the coding is the one oracle/deltacode.py defines and this file reproduces byte for byte:
//   base selection  head(i,b) = first first-occurrence chunk of the bucket (band b, key) chunk i falls in; votes(i,j) = bands
//                   whose head is j < i; root(i) = first occurrence with no j reaching min_votes;
//                   base(i) = best-voted root (ties to the smaller id)  -> bases never chain
//   delta           op* ; op = varint(len << 1 | kind) ; ADD: len literals ; COPY: varint(zigzag(q - expect))
//   encoder         H[h] = smallest base position per 14-bit hash of 8 bytes; greedy walk over the seeds of the
//                   target with backward + forward extension; kept iff 5 * bytes <= target length
// The per-position work (index build, seed test) is data-parallel over the CTA; the greedy walk is a serial
// chain of (find next seed, extend) steps, each of which one warp does cooperatively.
#include "ctx.cuh"

namespace {

constexpr uint32_t DELTA_MAX = 32768;
constexpr int DHB = 14;
constexpr uint32_t DEMPTY = 0xFFFFu;
constexpr uint64_t DMUL = 0x9E3779B97F4A7C15ull;
constexpr int DT = 256;  // threads per encode CTA

// ---- base selection ------------------------------------------------------------------------------------

// Sorted (band, key, id) triples: the head of a bucket is its first triple.  Gallop back to the bucket start.
__global__ void heads_kernel(const uint32_t* __restrict__ band, const uint64_t* __restrict__ key,
                             const uint64_t* __restrict__ id, uint64_t n_tr, uint32_t bands, uint64_t id_base,
                             const uint8_t* __restrict__ is_first, uint32_t* __restrict__ heads) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_tr) return;
    const uint32_t b = band[t];
    const uint64_t k = key[t];
    auto same = [&](int64_t c) { return band[c] == b && key[c] == k; };
    int64_t start = t;
    if (t > 0 && same(t - 1)) {
        int64_t hi = t - 1, lo, step = 1;
        for (;;) {
            if (hi < step) { lo = -1; break; }
            const int64_t c = hi - step;
            if (same(c)) { hi = c; step <<= 1; } else { lo = c; break; }
        }
        while (hi - lo > 1) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (same(mid)) hi = mid; else lo = mid;
        }
        start = hi;
    }
    // only first occurrences are ever inserted into the index (README.md:1553-1556): skip duplicates at the front of
    // the bucket (in one stream a duplicate is preceded by its first occurrence, so this loop does not run; it does on
    // a shard whose duplicates have their first occurrence elsewhere).  No first occurrence at all: the chunk itself.
    int64_t hd = start;
    if (is_first)
        while (hd < t && !is_first[id[hd] - id_base]) hd++;
    heads[(id[t] - id_base) * bands + b] = (uint32_t)(id[hd] - id_base);
}

// pass 0: root flags; pass 1: bases.  One warp per chunk, one lane per band (bands <= 32).  Chunk i of this call has id
// id_base + i and `heads` hold ids in the same space (single GPU: id_base 0; sharded: global ids of the first occurrences
// of the whole stream, where root_in covers ALL of them and root_out only this rank's).
template <int PASS>
__global__ void __launch_bounds__(256) votes_kernel(const uint32_t* __restrict__ heads, uint64_t n, uint32_t bands,
                                                    const uint8_t* __restrict__ is_first, uint32_t min_votes, uint64_t id_base,
                                                    uint8_t* __restrict__ root_out, const uint8_t* __restrict__ root_in,
                                                    int64_t* __restrict__ base) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (i >= n) return;
    const bool first = is_first ? is_first[i] != 0 : true;
    const uint32_t h = lane < bands ? heads[i * bands + lane] : 0xFFFFFFFFu;  // idle lanes: a value no chunk has
    const bool earlier = (uint64_t)h < id_base + i;
    uint32_t votes = __popc(__match_any_sync(0xFFFFFFFFu, h));
    if (!earlier) votes = 0;
    if (PASS == 0) {
        const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, votes);
        if (lane == 0) root_out[i] = first && best < min_votes;
    } else {
        const bool ok = earlier && votes >= min_votes && root_in[earlier ? h : 0];
        // most votes, then the smaller id
        const unsigned long long score = ok ? ((unsigned long long)votes << 32) | (0xFFFFFFFFu - h) : 0ull;
        unsigned long long best = score;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            best = v > best ? v : best;
        }
        if (lane == 0) base[i] = (first && best) ? (int64_t)(0xFFFFFFFFu - (uint32_t)best) : -1;
    }
}

// ---- encoder ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t ld64u(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p & ~(uintptr_t)7;
    const uint32_t sh = ((uint32_t)(uintptr_t)p & 7) * 8;
    const uint64_t lo = *reinterpret_cast<const uint64_t*>(a);
    if (sh == 0) return lo;
    const uint64_t hi = *reinterpret_cast<const uint64_t*>(a + 8);
    return (lo >> sh) | (hi << (64 - sh));
}
__device__ __forceinline__ uint32_t dhash(uint64_t w) { return (uint32_t)((w * DMUL) >> (64 - DHB)); }

__device__ __forceinline__ uint32_t varint_len(uint32_t v) { return v < 128 ? 1 : v < 16384 ? 2 : v < 2097152 ? 3 : 4; }
__device__ __forceinline__ uint8_t* put_varint(uint8_t* o, uint32_t v) {
    while (v >= 128) { *o++ = (uint8_t)(v | 0x80); v >>= 7; }
    *o++ = (uint8_t)v;
    return o;
}

// cap[i] = floor(len/5) for chunks that have a candidate base and fit the size limits, else 0; candidates are
// appended to `list` (processing order does not influence any output).
// A base index j >= n names external base j - n: ext_data[ext_off[j - n] : ext_off[j - n + 1]) (a chunk of another shard,
// fetched before the call); external bases precede their targets in the stream by construction.
__global__ void delta_plan_kernel(uint64_t start0, const uint64_t* __restrict__ cuts, uint64_t n, int64_t* __restrict__ base,
                                  const uint64_t* __restrict__ ext_off, uint64_t n_ext,
                                  uint64_t* __restrict__ cap, uint32_t* __restrict__ list, uint32_t* __restrict__ n_list) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t c = 0;
    const int64_t j = base[i];
    if (j >= 0) {
        const uint64_t len = cuts[i] - (i ? cuts[i - 1] : start0);
        const bool ext = (uint64_t)j >= n;
        uint64_t bl = DELTA_MAX + 1;
        if (!ext) bl = cuts[j] - (j ? cuts[j - 1] : start0);
        else if ((uint64_t)j - n < n_ext) bl = ext_off[j - n + 1] - ext_off[j - n];
        if ((ext || (uint64_t)j < i) && len <= DELTA_MAX && bl <= DELTA_MAX && len >= 5) {
            c = len / 5;
            list[atomicAdd(n_list, 1u)] = (uint32_t)i;
        } else {
            base[i] = -1;
        }
    }
    cap[i] = c;
}

struct EncArgs {
    const uint8_t* data;
    uint64_t start0;
    const uint64_t* cuts;
    int64_t* base;
    const uint32_t* list;
    const uint32_t* n_list;
    const uint64_t* slot;  // exclusive scan of cap
    uint8_t* stage;
    uint64_t* size;        // bytes of the kept delta (0: none)
    uint64_t n;            // chunks of this call: base indices >= n are external
    const uint8_t* ext_data;
    const uint64_t* ext_off;
};

__global__ void __launch_bounds__(DT) delta_encode_kernel(EncArgs a) {
    extern __shared__ uint32_t dsm[];
    uint32_t* H = dsm;                      // 1 << DHB entries
    uint32_t* seedbits = dsm + (1 << DHB);  // DELTA_MAX / 32 words
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_list = *a.n_list;
    for (uint32_t w = blockIdx.x; w < n_list; w += gridDim.x) {
        const uint64_t i = a.list[w];
        const uint64_t j = (uint64_t)a.base[i];
        const uint64_t t0 = i ? a.cuts[i - 1] : a.start0;
        const uint32_t n = (uint32_t)(a.cuts[i] - t0);
        const uint8_t* T = a.data + t0;
        const uint8_t* B;
        uint32_t nb;
        if (j >= a.n) {
            B = a.ext_data + a.ext_off[j - a.n];
            nb = (uint32_t)(a.ext_off[j - a.n + 1] - a.ext_off[j - a.n]);
        } else {
            const uint64_t b0 = j ? a.cuts[j - 1] : a.start0;
            B = a.data + b0;
            nb = (uint32_t)(a.cuts[j] - b0);
        }
        __syncthreads();  // the previous pair's walk is done with H / seedbits
        for (uint32_t k = tid; k < (1u << DHB); k += DT) H[k] = DEMPTY;
        __syncthreads();
        if (nb >= 8)
            for (uint32_t q = tid; q + 8 <= nb; q += DT) atomicMin(&H[dhash(ld64u(B + q))], q);
        __syncthreads();
        const uint32_t n_words = (n + 31) >> 5;
        for (uint32_t sw = warp; sw < n_words; sw += DT / 32) {
            const uint32_t s = sw * 32 + lane;
            bool ok = false;
            if (s + 8 <= n) {
                const uint64_t x = ld64u(T + s);
                const uint32_t q = H[dhash(x)];
                ok = q != DEMPTY && ld64u(B + q) == x;
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
            if (lane == 0) seedbits[sw] = m;
        }
        __syncthreads();
        if (warp != 0) continue;
        // ---- the greedy walk: warp-uniform state, lanes share each step ----
        const uint32_t cap = n / 5;
        uint8_t* out = a.stage + a.slot[i];
        uint32_t p = 0, o = 0, expect = 0;
        bool keep = true;
        while (keep) {
            // first seed at or after p
            uint32_t s = 0xFFFFFFFFu;
            {
                uint32_t wi = p >> 5;
                uint32_t m = wi < n_words ? (seedbits[wi] & (0xFFFFFFFFu << (p & 31))) : 0u;
                if (m) s = wi * 32 + (__ffs(m) - 1);
                else {
                    for (wi += 1; wi < n_words; wi += 32) {
                        const uint32_t mm = wi + lane < n_words ? seedbits[wi + lane] : 0u;
                        const uint32_t any = __ballot_sync(0xFFFFFFFFu, mm != 0);
                        if (any) {
                            const int src = __ffs(any) - 1;
                            const uint32_t mw = __shfl_sync(0xFFFFFFFFu, mm, src);
                            s = (wi + src) * 32 + (__ffs(mw) - 1);
                            break;
                        }
                    }
                }
            }
            if (s == 0xFFFFFFFFu) break;
            uint32_t q = H[dhash(ld64u(T + s))];
            // backwards: not past p, not past the start of the base
            {
                const uint32_t maxback = min(s - p, q);
                uint32_t k = 0;
                while (k < maxback) {
                    const uint32_t kk = k + lane;
                    const bool eq = kk < maxback && T[s - 1 - kk] == B[q - 1 - kk];
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, !eq);
                    if (m) { k += __ffs(m) - 1; break; }
                    k += 32;
                }
                k = min(k, maxback);
                s -= k;
                q -= k;
            }
            // forwards from s + 8, eight bytes per lane per step
            uint32_t L = 8;
            {
                const uint32_t lim = min(n - s, nb - q);
                while (L < lim) {
                    const uint32_t off = L + lane * 8;
                    uint32_t good = 0;  // agreeing bytes of this lane's eight
                    if (off < lim) {
                        const uint64_t x = ld64u(T + s + off) ^ ld64u(B + q + off);
                        good = x ? (uint32_t)(__ffsll((long long)x) - 1) >> 3 : 8u;
                        good = min(good, lim - off);
                    }
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, good < 8);
                    if (m) {
                        const int src = __ffs(m) - 1;
                        L += src * 8 + __shfl_sync(0xFFFFFFFFu, good, src);
                        break;
                    }
                    L += 256;
                }
                L = min(L, lim);
            }
            const uint32_t nlit = s - p;
            const uint32_t d = q >= expect ? (q - expect) << 1 : ((expect - q) << 1) - 1;
            const uint32_t need = (nlit ? varint_len(nlit << 1) + nlit : 0) + varint_len((L << 1) | 1) + varint_len(d);
            if (o + need > cap) { keep = false; break; }
            if (nlit) {
                const uint32_t hl = varint_len(nlit << 1);
                if (lane == 0) put_varint(out + o, nlit << 1);
                for (uint32_t k = lane; k < nlit; k += 32) out[o + hl + k] = T[p + k];
                o += hl + nlit;
            }
            if (lane == 0) put_varint(put_varint(out + o, (L << 1) | 1), d);
            o += varint_len((L << 1) | 1) + varint_len(d);
            expect = q + L;
            p = s + L;
        }
        if (keep && p < n) {
            const uint32_t nlit = n - p, hl = varint_len(nlit << 1);
            if (o + hl + nlit > cap) keep = false;
            else {
                if (lane == 0) put_varint(out + o, nlit << 1);
                for (uint32_t k = lane; k < nlit; k += 32) out[o + hl + k] = T[p + k];
                o += hl + nlit;
            }
        }
        if (lane == 0) {
            a.size[i] = keep ? o : 0;
            if (!keep) a.base[i] = -1;
        }
    }
}

// ---- decoder ---------------------------------------------------------------------------------------------

struct ApplyArgs {
    const uint8_t* delta;
    const uint64_t* delta_off;
    uint64_t m;
    const uint8_t* base;
    uint64_t base_bytes;
    const uint64_t* base_off;
    const uint32_t* base_len;
    uint8_t* out;
    const uint64_t* out_off;
    uint32_t* status;
    unsigned long long* bad;
};

// One warp per delta; the op stream is serial, the copies are shared by the lanes.
__global__ void __launch_bounds__(256) delta_apply_kernel(ApplyArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < a.m; j += warps) {
        const uint8_t* D = a.delta + a.delta_off[j];
        const uint64_t dl64 = a.delta_off[j + 1] - a.delta_off[j];
        const uint8_t* B = a.base + a.base_off[j];
        const uint32_t nb = a.base_len[j];
        uint8_t* O = a.out + a.out_off[j];
        const uint64_t n64 = a.out_off[j + 1] - a.out_off[j];
        uint32_t st = 0;
        if (dl64 > 0xFFFFFFFFull || n64 > 0xFFFFFFFFull) st = 2;   // (also offsets that run backwards)
        if (a.base_off[j] > a.base_bytes || nb > a.base_bytes - a.base_off[j]) st = 3;   // base range outside d_base
        const uint32_t dl = (uint32_t)dl64, n = (uint32_t)n64;
        uint32_t i = 0, o = 0, expect = 0;
        // varint at D[i]: lanes 0..4 fetch one byte each
        auto rd = [&](uint32_t& v) -> bool {
            const uint32_t b = (lane < 5 && i + lane < dl) ? D[i + lane] : 0u;
            const uint32_t cont = __ballot_sync(0xFFFFFFFFu, (b & 0x80) != 0) & 31u;
            const uint32_t len = __ffs(~cont);  // first byte without the continuation bit, 1-based
            if (len > 5 || i + len > dl) return false;
            if (len == 5 && __shfl_sync(0xFFFFFFFFu, b, 4) > 0x0F) return false;  // more than 32 bits
            v = __reduce_or_sync(0xFFFFFFFFu, lane < len ? (b & 0x7F) << (7 * lane) : 0u);
            i += len;
            return true;
        };
        while (!st && o < n) {
            uint32_t tag, z;
            if (!rd(tag)) { st = 1; break; }
            const uint32_t ln = tag >> 1;
            if (ln == 0 || ln > n - o) { st = 2; break; }
            if (tag & 1) {
                if (!rd(z)) { st = 1; break; }
                const int64_t q = (int64_t)expect + ((z & 1) ? -(int64_t)((z + 1) >> 1) : (int64_t)(z >> 1));
                if (q < 0 || q + ln > (int64_t)nb) { st = 3; break; }
                for (uint32_t k = lane; k < ln; k += 32) O[o + k] = B[q + k];
                expect = (uint32_t)q + ln;
            } else {
                if (ln > dl - i) { st = 4; break; }
                for (uint32_t k = lane; k < ln; k += 32) O[o + k] = D[i + k];
                i += ln;
            }
            o += ln;
        }
        if (!st && i != dl) st = 5;
        if (lane == 0) {
            a.status[j] = st;
            if (st) atomicAdd(a.bad, 1ull);
        }
    }
}

}  // namespace

HMSE_API int hmse_delta_bases(hmse_ctx* ctx, const uint32_t* d_band, const uint64_t* d_key, const uint64_t* d_id, uint64_t n,
                              uint32_t bands, uint64_t id_base, const uint8_t* d_is_first, uint32_t min_votes,
                              int64_t* d_base, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (n == 0) return HMSE_OK;
    if (!d_band || !d_key || !d_id || !d_is_first || !d_base) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_bases: null pointer");
    if (bands == 0 || bands > 32) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_bases: bands must be 1..32");
    if (n > 0xFFFFFFFEull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_bases: n exceeds 2^32 - 2");
    if (min_votes == 0) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_bases: min_votes must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    // misc: [heads u32 n*bands][root u8 n]
    HMSE_SCRATCH(ctx, heads, uint32_t*, SLOT_DELTA_HEADS, n * bands * 4 + n + 16);
    uint8_t* root = reinterpret_cast<uint8_t*>(heads + n * bands);
    HT_BEGIN(ctx, HT_DELTA, st);
    KL(ctx);
    heads_kernel<<<(unsigned)div_up64(n * bands, 256), 256, 0, st>>>(d_band, d_key, d_id, n * bands, bands, id_base, d_is_first,
                                                                     heads);
    KL(ctx);
    votes_kernel<0><<<(unsigned)div_up64(n * 32, 256), 256, 0, st>>>(heads, n, bands, d_is_first, min_votes, 0, root, root, d_base);
    KL(ctx);
    votes_kernel<1><<<(unsigned)div_up64(n * 32, 256), 256, 0, st>>>(heads, n, bands, d_is_first, min_votes, 0, root, root, d_base);
    HMSE_LAUNCH_CHECK(ctx);
    HT_END(ctx, HT_DELTA, st);
    return HMSE_OK;
}

HMSE_API int hmse_delta_encode(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, uint64_t n,
                               int64_t* d_base, uint8_t* d_out, uint64_t out_cap, uint64_t* d_offsets, uint64_t* total,
                               void* stream) {
    return hmse_delta_encode_ext(ctx, d_data, start0, d_cuts, n, d_base, nullptr, nullptr, 0, d_out, out_cap, d_offsets, total, stream);
}

HMSE_API int hmse_delta_heads(hmse_ctx* ctx, const uint32_t* d_band, const uint64_t* d_key, const uint64_t* d_id, uint64_t n,
                              uint32_t bands, uint64_t id_base, const uint8_t* d_is_first, uint32_t* d_heads, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (n == 0 || bands == 0) return HMSE_OK;
    if (!d_band || !d_key || !d_id || !d_heads) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_heads: null pointer");
    if (n > 0xFFFFFFFEull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_heads: n exceeds 2^32 - 2");
    KL(ctx);
    heads_kernel<<<(unsigned)div_up64(n * bands, 256), 256, 0, (cudaStream_t)stream>>>(d_band, d_key, d_id, n * bands, bands, id_base,
                                                                                      d_is_first, d_heads);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}

HMSE_API int hmse_delta_votes(hmse_ctx* ctx, const uint32_t* d_heads, uint64_t n, uint32_t bands, uint64_t id_base,
                              uint32_t min_votes, int pass, uint8_t* d_root_local, const uint8_t* d_root_all, int64_t* d_base,
                              void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (n == 0) return HMSE_OK;
    if (bands == 0 || bands > 32 || min_votes == 0) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_votes: bands must be 1..32, min_votes >= 1");
    if (!d_heads || (pass == 0 && !d_root_local) || (pass == 1 && (!d_root_all || !d_base)) || (pass != 0 && pass != 1))
        HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_votes: null pointer or bad pass");
    if (id_base + n > 0xFFFFFFFEull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_votes: ids exceed 2^32 - 2");
    cudaStream_t st = (cudaStream_t)stream;
    KL(ctx);
    if (pass == 0)
        votes_kernel<0><<<(unsigned)div_up64(n * 32, 256), 256, 0, st>>>(d_heads, n, bands, nullptr, min_votes, id_base, d_root_local,
                                                                         nullptr, nullptr);
    else
        votes_kernel<1><<<(unsigned)div_up64(n * 32, 256), 256, 0, st>>>(d_heads, n, bands, nullptr, min_votes, id_base, nullptr,
                                                                         d_root_all, d_base);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}

HMSE_API int hmse_delta_encode_ext(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts, uint64_t n,
                                   int64_t* d_base, const uint8_t* d_ext_data, const uint64_t* d_ext_off, uint64_t n_ext,
                                   uint8_t* d_out, uint64_t out_cap, uint64_t* d_offsets, uint64_t* total, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (n_ext && (!d_ext_data || !d_ext_off)) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode_ext: null external bases");
    if (n_ext && ((uintptr_t)d_ext_data & 7)) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode_ext: d_ext_data must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (total) *total = 0;
    if (!d_offsets) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode: null d_offsets");
    if (n == 0) {
        HMSE_CUDA(ctx, cudaMemsetAsync(d_offsets, 0, 8, st));
        return HMSE_OK;
    }
    if (!d_data || !d_cuts || !d_base || !total) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode: null pointer");
    if (n > 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode: n exceeds 2^32");
    if ((uintptr_t)d_data & 15) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode: d_data must be 16-byte aligned");
    // misc: [cap/slot u64 n+1][size u64 n+1][list u32 n][n_list u32, pad]
    HMSE_SCRATCH(ctx, misc, uint64_t*, SLOT_DELTA_MISC, (2 * (n + 1)) * 8 + n * 4 + 16);
    uint64_t* slot = misc;
    uint64_t* size = misc + (n + 1);
    uint32_t* list = reinterpret_cast<uint32_t*>(misc + 2 * (n + 1));
    uint32_t* n_list = list + n;
    HT_BEGIN(ctx, HT_DELTA, st);
    HMSE_CUDA(ctx, cudaMemsetAsync(n_list, 0, 4, st));
    HMSE_CUDA(ctx, cudaMemsetAsync(size, 0, (n + 1) * 8, st));
    KL(ctx);
    delta_plan_kernel<<<(unsigned)div_up64(n, 256), 256, 0, st>>>(start0, d_cuts, n, d_base, d_ext_off, n_ext, slot, list, n_list);
    HMSE_LAUNCH_CHECK(ctx);
    if (int rc = hmse_exclusive_scan_u64(ctx, slot, slot, n, slot + n, st)) return rc;
    if (int rc = hmse_mail(ctx, 0, slot + n, 2, st)) return rc;
    if (int rc = hmse_mail(ctx, 2, n_list, 1, st)) return rc;
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t stage_bytes = ctx->pinned[0];
    const uint32_t cands = (uint32_t)ctx->pinned[1];
    if (cands) {
        HMSE_SCRATCH(ctx, stage, uint8_t*, SLOT_DELTA_STAGE, stage_bytes + 16);
        EncArgs ea{d_data, start0, d_cuts, d_base, list, n_list, slot, stage, size, n, d_ext_data, d_ext_off};
        const uint32_t grid = cands < (uint32_t)ctx->sm_count * 3 ? cands : (uint32_t)ctx->sm_count * 3;
        const size_t smem = ((1u << DHB) + DELTA_MAX / 32) * 4;
        HMSE_CUDA(ctx, cudaFuncSetAttribute(delta_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KL(ctx);
        delta_encode_kernel<<<grid, DT, smem, st>>>(ea);
        HMSE_LAUNCH_CHECK(ctx);
    }
    if (int rc = hmse_exclusive_scan_u64(ctx, size, d_offsets, n, d_offsets + n, st)) return rc;
    if (int rc = hmse_mail(ctx, 0, d_offsets + n, 2, st)) return rc;
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    *total = ctx->pinned[0];
    if (*total > out_cap) HMSE_FAIL(ctx, HMSE_E_CAPACITY, "hmse_delta_encode: out_cap %llu < %llu",
                                    (unsigned long long)out_cap, (unsigned long long)*total);
    if (*total) {
        if (!d_out) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_encode: null d_out");
        uint8_t* stage = (uint8_t*)ctx->slot[SLOT_DELTA_STAGE];
        if (int rc = hmse_segment_copy(ctx, stage, slot, d_out, d_offsets, n, st)) return rc;
    }
    HT_END(ctx, HT_DELTA, st);
    return HMSE_OK;
}

HMSE_API int hmse_delta_apply(hmse_ctx* ctx, const uint8_t* d_delta, const uint64_t* d_delta_off, uint64_t m,
                              const uint8_t* d_base, uint64_t base_bytes, const uint64_t* d_base_off, const uint32_t* d_base_len,
                              uint8_t* d_out, const uint64_t* d_out_off, uint32_t* d_status, uint64_t* n_bad, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_bad) *n_bad = 0;
    if (m == 0) return HMSE_OK;
    if (!d_delta || !d_delta_off || !d_base || !d_base_off || !d_base_len || !d_out || !d_out_off || !d_status)
        HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_delta_apply: null pointer");
    HMSE_SCRATCH(ctx, bad, unsigned long long*, SLOT_DELTA_BAD, 16);
    HMSE_CUDA(ctx, cudaMemsetAsync(bad, 0, 8, st));
    ApplyArgs aa{d_delta, d_delta_off, m, d_base, base_bytes, d_base_off, d_base_len, d_out, d_out_off, d_status, bad};
    const uint64_t want = div_up64(m, 8), cap = (uint64_t)ctx->sm_count * 8;
    KL(ctx);
    delta_apply_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(aa);
    HMSE_LAUNCH_CHECK(ctx);
    if (n_bad) {
        if (int rc = hmse_mail(ctx, 0, bad, 2, st)) return rc;
        HMSE_CUDA(ctx, cudaStreamSynchronize(st));
        *n_bad = ctx->pinned[0];
    }
    return HMSE_OK;
}
