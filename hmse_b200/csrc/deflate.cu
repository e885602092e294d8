// deflate.cu - L1 per-chunk DEFLATE with a preset dictionary (spec: README.md:288, 1159-1198;
// the skeleton's mz_deflateInit2(..., 15, ...) / mz_deflate(FINISH) at README.md:2374, 2378;
// resolved to one RFC 1950 stream per chunk with FDICT in SURVEY.md §0.2 C5).
//
// One CTA compresses one chunk at a time (persistent CTAs pull chunks from a counter, two size
// classes so small chunks get 3 CTAs per SM).  Unlike zlib's serial hash-chain walk, the best
// match of EVERY position is found in parallel, then the zlib level-6 lazy rule is applied as a
// pure function next(p) so the parse becomes chain following:
//   P0 stage chunk into shared memory          P5 per-32-byte-range backward DP of chain exits
//   P1 histogram of 4-byte hashes (13 bits)    P6 hop the true chain across ranges
//   P2 scan -> bucket ends                     P7 symbol histograms of the visited tokens
//   P3 scatter positions into buckets          P8 Huffman lengths/codes + dynamic header
//   P4 match search: own buckets in shared,    P9 bit counts per range + block scan
//      dictionary buckets (host-built index,   P10 parallel bit emission into the stage slot
//      L1/L2 resident) in global memory        (stored / fixed / dynamic chosen like zlib)
// A pack kernel then compacts the per-chunk stage slots into the caller's blob.
// Output is compared with zlib by inflate-equality and total size only (never byte for byte).
#include <stdlib.h>

#include "ctx.cuh"
#include "deflate_core.h"

using namespace dfl;

namespace {

constexpr int OWN_CAP = 48;    // own-chunk candidates examined per position
constexpr int DICT_CAP = 32;   // dictionary candidates examined per position
constexpr uint32_t DICT_MAX = 32768;
constexpr uint32_t NMAX_SMALL = 12288, NMAX_LARGE = 32768;
constexpr int T_SMALL = 256, T_LARGE = 512;

// Device image of the dictionary index (SLOT_DEFLATE_DICT), built on the host once per dictionary.
struct DictDev {
    uint8_t bytes[DICT_MAX + 16];
    uint16_t boff[NBUCKET + 8];   // bucket h = [boff[h], boff[h+1]) in sorted order
    uint16_t sorted[DICT_MAX];    // positions, nearest (largest) first inside a bucket
    uint32_t first4[DICT_MAX];    // le32 at the position, same order as sorted[]
};

struct DeflArgs {
    const uint8_t* data;
    uint64_t start0;
    const uint64_t* cuts;
    const uint64_t* select;  // may be null
    const DictDev* dict;
    uint32_t dict_len, dict_adler;
    int level;
    const uint32_t* list;    // selection slots of this size class
    const uint32_t* list_n;
    uint32_t nmax;
    uint8_t* stage;
    const uint64_t* slot_off;
    uint64_t* sizes;
    uint32_t* match;         // per-CTA scratch, nmax words each
    unsigned int* counter;
};

// Per-phase cycle counters (thread 0 of every CTA, summed over chunks); read by
// hmse_debug_deflate_prof.  A dozen clock reads per chunk: negligible.
__device__ unsigned long long g_prof[16];
#define PROF(i)                                             \
    if (t == 0) {                                           \
        const long long now__ = clock64();                  \
        atomicAdd(&g_prof[i], (unsigned long long)(now__ - tprev)); \
        tprev = now__;                                      \
    }

struct Small {  // fixed-size shared state
    uint32_t hist_lit[288];
    uint32_t hist_dist[32];
    uint32_t code_lit[288];
    uint32_t code_dist[32];
    uint32_t warp_tmp[40];
    uint8_t hdr[640];
    uint32_t hdr_bits, mode, total_bits, job, n_used;
    uint32_t adler_a, adler_b;
};

__device__ __forceinline__ uint32_t ld32u(const uint32_t* w, uint32_t off) {
    const uint32_t i = off >> 2;
    return __funnelshift_r(w[i], w[i + 1], (off & 3) * 8);
}
__device__ __forceinline__ uint32_t ldg32u(const uint8_t* base, uint32_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (off >> 2);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), (off & 3) * 8);
}

// Exclusive scan of one u32 per thread; *total = block sum (same value in every thread).
// tmp = 33 words of shared memory.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* tmp, uint32_t* total) {
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    __syncthreads();
    if (lane == 31) tmp[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < nw ? tmp[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (unsigned)o) si += t;
        }
        tmp[lane] = si - s;
        if (lane == 31) tmp[32] = si;  // zeros past nw: lane 31 holds the grand total
    }
    __syncthreads();
    const uint32_t r = inc - v + tmp[w];
    *total = tmp[32];
    __syncthreads();
    return r;
}

// Token at p given the (lazy-resolved) match word: returns step, fills symbol info.
struct Tok {
    uint32_t step, lsym, lbits, lval, dsym, dbits, dval;
    bool is_match;
};
__device__ __forceinline__ Tok token_at(uint32_t mw, uint32_t byte) {
    Tok t;
    const uint32_t L = mw >> 16;
    if (L == 0 || (mw & 0x8000u)) {
        t.is_match = false;
        t.step = 1;
        t.lsym = byte;
        t.lbits = t.lval = t.dsym = t.dbits = t.dval = 0;
    } else {
        t.is_match = true;
        t.step = L;
        len_sym(L, t.lsym, t.lbits, t.lval);
        dist_sym((mw & 0x7fffu) + 1, t.dsym, t.dbits, t.dval);
    }
    return t;
}

struct Emitter {
    uint32_t* out;      // 4-byte aligned start of the deflate bit stream
    uint64_t acc;
    uint32_t accbits, wi;
    bool first;
    __device__ __forceinline__ void begin(uint32_t* o, uint32_t bitoff) {
        out = o;
        wi = bitoff >> 5;
        accbits = bitoff & 31;
        acc = 0;
        first = true;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t nb) {
        acc |= (uint64_t)v << accbits;
        accbits += nb;
        if (accbits >= 32) {
            if (first) atomicOr(out + wi, (uint32_t)acc);  // shared with the previous writer
            else out[wi] = (uint32_t)acc;
            first = false;
            acc >>= 32;
            accbits -= 32;
            wi++;
        }
    }
    __device__ __forceinline__ void end() {
        if (accbits) atomicOr(out + wi, (uint32_t)acc);
    }
};

__global__ void __launch_bounds__(T_LARGE) deflate_kernel(DeflArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t nmax = a.nmax;
    // layout
    uint32_t* s_data32 = reinterpret_cast<uint32_t*>(smem);                     // nmax + 16 bytes
    const uint8_t* s_data = smem;
    uint16_t* s_sorted = reinterpret_cast<uint16_t*>(smem + nmax + 16);        // nmax u16 (later: exits)
    uint32_t* s_cnt32 = reinterpret_cast<uint32_t*>(smem + nmax + 16 + 2 * (size_t)nmax);  // NBUCKET/2 + 4 words
    uint16_t* s_cnt16 = reinterpret_cast<uint16_t*>(s_cnt32);
    uint8_t* s_entry = reinterpret_cast<uint8_t*>(s_cnt32 + NBUCKET / 2 + 4);   // nmax/32 bytes
    Small* sm = reinterpret_cast<Small*>(s_entry + nmax / 32);
    // Huffman scratch aliases the bucket table (dead after P4)
    HuffWork* hw = reinterpret_cast<HuffWork*>(s_cnt32);
    DynHeader* dh = reinterpret_cast<DynHeader*>(reinterpret_cast<uint8_t*>(s_cnt32) + ((sizeof(HuffWork) + 15) & ~15u));
    uint16_t* s_exit = s_sorted;
    uint32_t* g_match = a.match + (size_t)blockIdx.x * nmax;
    const uint32_t list_n = *a.list_n;

    for (;;) {
        __syncthreads();
        if (t == 0) sm->job = atomicAdd(a.counter, 1u);
        __syncthreads();
        const uint32_t job = sm->job;
        if (job >= list_n) break;
        long long tprev = clock64();
        const uint32_t k = a.list[job];
        const uint64_t j = a.select ? a.select[k] : (uint64_t)k;
        const uint64_t cs = j ? a.cuts[j - 1] : a.start0;
        const uint32_t n = (uint32_t)(a.cuts[j] - cs);
        const uint8_t* src = a.data + cs;
        uint8_t* slot = a.stage + a.slot_off[k];
        const uint32_t hdr_len = a.dict_len ? 6 : 2;
        uint32_t* out32 = reinterpret_cast<uint32_t*>(slot + 2 + hdr_len);  // slot is 16-aligned: 4 or 8

        // ---- P0: stage the chunk (byte-unaligned source -> aligned words), clear tables -----
        {
            const uint32_t kmis = (uint32_t)((uintptr_t)src & 3);
            const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(src - kmis);
            const uint32_t nw = (n + 3) >> 2;
            for (uint32_t i = t; i < nw + 4 && i < (nmax + 16) / 4; i += T) {
                uint32_t v = 0;
                if (i < nw) {
                    uint32_t lo = __ldg(wsrc + i);
                    uint32_t hi = kmis ? __ldg(wsrc + i + 1) : 0u;
                    v = __funnelshift_r(lo, hi, kmis * 8);
                    const uint32_t rem = n - 4 * i;  // bytes of this word inside the chunk
                    if (rem < 4) v &= (1u << (8 * rem)) - 1;
                }
                s_data32[i] = v;
            }
            for (uint32_t i = t; i < NBUCKET / 2 + 4; i += T) s_cnt32[i] = 0;
            for (uint32_t i = t; i < 288; i += T) sm->hist_lit[i] = 0;
            if (t < 32) sm->hist_dist[t] = 0;
            for (uint32_t i = t; i < 160; i += T) reinterpret_cast<uint32_t*>(sm->hdr)[i] = 0;
        }
        __syncthreads();
        PROF(0)
        const uint32_t nh = n >= 4 ? n - 3 : 0;  // hashed positions

        // ---- Adler-32 of the chunk (parallel partial sums, n <= 32768) --------------------
        {
            uint32_t sa = 0, sb = 0;
            for (uint32_t p = t; p < n; p += T) {
                uint32_t b = s_data[p];
                sa += b;
                sb += (n - p) * b;  // <= 32768*255 per term, <= 128 terms per thread at T>=256
                if (sb >= 0x80000000u) sb %= 65521u;
            }
            sb %= 65521u;
            uint32_t tot_a, tot_b;
            block_excl_scan(sa, sm->warp_tmp, &tot_a);
            block_excl_scan(sb, sm->warp_tmp, &tot_b);
            if (t == 0) {
                sm->adler_a = (1u + tot_a) % 65521u;
                sm->adler_b = (n % 65521u + tot_b) % 65521u;
            }
        }

        PROF(1)
        const bool stored_only = a.level == 0;
        if (!stored_only) {
            // ---- P1: hash histogram --------------------------------------------------------
            for (uint32_t p = t; p < nh; p += T) {
                const uint32_t h = hash4(ld32u(s_data32, p));
                atomicAdd(&s_cnt32[h >> 1], 1u << (16 * (h & 1)));
            }
            __syncthreads();
            PROF(2)
            // ---- P2: inclusive scan over buckets -> bucket ends ------------------------------
            {
                const uint32_t per = NBUCKET / T;  // buckets per thread (T divides NBUCKET)
                uint32_t sum = 0;
                for (uint32_t i = 0; i < per; i++) sum += s_cnt16[t * per + i];
                uint32_t tot;
                uint32_t run = block_excl_scan(sum, sm->warp_tmp, &tot);
                for (uint32_t i = 0; i < per; i++) {
                    run += s_cnt16[t * per + i];
                    s_cnt16[t * per + i] = (uint16_t)run;
                }
                if (t == 0) s_cnt16[NBUCKET] = (uint16_t)nh;
            }
            __syncthreads();
            PROF(3)
            // ---- P3: scatter (ends count down to starts) ------------------------------------
            for (uint32_t p = t; p < nh; p += T) {
                const uint32_t h = hash4(ld32u(s_data32, p));
                const uint32_t sh = 16 * (h & 1);
                const uint32_t old = atomicSub(&s_cnt32[h >> 1], 1u << sh);
                s_sorted[((old >> sh) & 0xffffu) - 1] = (uint16_t)p;
            }
            __syncthreads();
            PROF(4)
            // ---- P4: best match of every position -------------------------------------------
            for (uint32_t p = t; p < n; p += T) {
                uint32_t best = 3, bdist = 0;
                if (p < nh) {
                    const uint32_t v = ld32u(s_data32, p);
                    const uint32_t h = hash4(v);
                    const uint32_t maxl = n - p < (uint32_t)MAX_MATCH ? n - p : (uint32_t)MAX_MATCH;
                    const uint32_t b0 = s_cnt16[h], b1 = s_cnt16[h + 1];
                    int ex = 0;
                    for (uint32_t i = b0; i < b1 && ex < OWN_CAP; i++) {
                        const uint32_t q = s_sorted[i];
                        if (q >= p) continue;
                        ex++;
                        if (ld32u(s_data32, q) != v) continue;
                        uint32_t l = 4;
                        while (l < maxl) {
                            const uint32_t x = ld32u(s_data32, p + l) ^ ld32u(s_data32, q + l);
                            if (x) {
                                l += (uint32_t)(__ffs((int)x) - 1) >> 3;
                                break;
                            }
                            l += 4;
                        }
                        if (l > maxl) l = maxl;
                        const uint32_t dist = p - q;
                        if (l > best || (l == best && dist < bdist)) {
                            best = l;
                            bdist = dist;
                        }
                        if (best >= (uint32_t)NICE_LENGTH || best == maxl) break;
                    }
                    if (a.dict_len && best < (uint32_t)NICE_LENGTH && best < maxl) {
                        const uint32_t d0 = __ldg(&a.dict->boff[h]), d1 = __ldg(&a.dict->boff[h + 1]);
                        int exd = 0;
                        for (uint32_t i = d0; i < d1 && exd < DICT_CAP; i++, exd++) {
                            if (__ldg(&a.dict->first4[i]) != v) continue;
                            const uint32_t jpos = __ldg(&a.dict->sorted[i]);
                            const uint32_t dist = p + a.dict_len - jpos;
                            if (dist > (uint32_t)WSIZE) continue;
                            uint32_t lim = a.dict_len - jpos;  // matches do not run from the dictionary into the chunk
                            if (lim > maxl) lim = maxl;
                            uint32_t l = 4;
                            while (l < lim) {
                                const uint32_t x = ld32u(s_data32, p + l) ^ ldg32u(a.dict->bytes, jpos + l);
                                if (x) {
                                    l += (uint32_t)(__ffs((int)x) - 1) >> 3;
                                    break;
                                }
                                l += 4;
                            }
                            if (l > lim) l = lim;
                            if (l > best) {  // equal length: the own-chunk match is nearer
                                best = l;
                                bdist = dist;
                            }
                            if (best >= (uint32_t)NICE_LENGTH || best == maxl) break;
                        }
                    }
                }
                g_match[p] = best >= 4 ? (best << 16) | (bdist - 1) : 0u;
            }
            __syncthreads();
            PROF(5)
        }

        // ranges of 32 positions, blocked over threads
        const uint32_t R = (n + 31) >> 5;
        const uint32_t rpt = (R + T - 1) / T;
        const uint32_t r0 = t * rpt, r1 = (r0 + rpt < R) ? r0 + rpt : R;

        if (!stored_only) {
            // ---- P5: backward DP: exit[p] = first chain position past p's range -------------
            for (uint32_t r = r0; r < r1; r++) {
                const uint32_t ps = r << 5, pe = (ps + 32 < n) ? ps + 32 : n;
                uint32_t nxt_len = pe < n ? (g_match[pe] >> 16) : 0;
                for (uint32_t p = pe; p-- > ps;) {
                    const uint32_t mw = g_match[p];
                    const uint32_t L = mw >> 16;
                    const bool lazy_lit = L != 0 && L < (uint32_t)MAX_LAZY && nxt_len > L;
                    if (lazy_lit) g_match[p] = mw | 0x8000u;
                    const uint32_t next = (L == 0 || lazy_lit) ? p + 1 : p + L;
                    s_exit[p] = (uint16_t)(next >= pe ? next : s_exit[next]);
                    nxt_len = L;
                }
            }
            for (uint32_t r = t; r < R; r += T) s_entry[r] = 0xFF;
            __syncthreads();
            PROF(6)
            // ---- P6: hop the true chain across ranges -----------------------------------------
            if (t == 0) {
                uint32_t p = 0;
                while (p < n) {
                    s_entry[p >> 5] = (uint8_t)(p & 31);
                    p = s_exit[p];
                }
            }
            __syncthreads();
            PROF(7)
            // ---- P7: symbol histograms ----------------------------------------------------------
            for (uint32_t r = r0; r < r1; r++) {
                const uint32_t e = s_entry[r];
                if (e == 0xFF) continue;
                const uint32_t pe = ((r << 5) + 32 < n) ? (r << 5) + 32 : n;
                for (uint32_t p = (r << 5) + e; p < pe;) {
                    const Tok tk = token_at(g_match[p], s_data[p]);
                    atomicAdd(&sm->hist_lit[tk.lsym], 1u);
                    if (tk.is_match) atomicAdd(&sm->hist_dist[tk.dsym], 1u);
                    p += tk.step;
                }
            }
            if (t == 0) atomicAdd(&sm->hist_lit[EOB], 1u);
            __syncthreads();
            PROF(8)
            // ---- P8: Huffman codes + header -------------------------------------------------------
            // parallel rank sort of the literal/length alphabet by (freq, symbol)
            if (t == 0) {
                uint32_t used = 0;
                for (int s = 0; s < NLIT; s++) used += sm->hist_lit[s] != 0;
                if (used < 2) sm->hist_lit[sm->hist_lit[0] ? 1 : 0] = 1;  // EOB is always used: force a second code
                sm->n_used = used < 2 ? 2 : used;
            }
            __syncthreads();
            for (uint32_t s = t; s < (uint32_t)NLIT; s += T) {
                const uint32_t f = sm->hist_lit[s];
                if (f) {
                    uint32_t rank = 0;
                    for (uint32_t o = 0; o < (uint32_t)NLIT; o++) {
                        const uint32_t g = sm->hist_lit[o];
                        rank += (g != 0) && (g < f || (g == f && o < s));
                    }
                    hw->w[rank] = f;
                    hw->order[rank] = (uint16_t)s;
                }
            }
            __syncthreads();
            if (t == 0) {
                build_lengths_sorted(*hw, (int)sm->n_used, 15, dh->lit_lens, NLIT);
                uint32_t df[32];
                for (int s = 0; s < 32; s++) df[s] = sm->hist_dist[s];
                build_lengths(df, NDIST, 15, dh->dist_lens, *hw);
                plan_header_from_lengths(*dh, *hw);
                uint64_t dyn_bits = dh->bits, fix_bits = 3;
                for (int s = 0; s < NLIT; s++) {
                    const uint32_t f = sm->hist_lit[s];
                    const int xb = s > 256 ? lsym_extra(s) : 0;
                    dyn_bits += (uint64_t)f * (dh->lit_lens[s] + xb);
                    fix_bits += (uint64_t)f * (fixed_lit_len(s) + xb);
                }
                for (int s = 0; s < NDIST; s++) {
                    dyn_bits += (uint64_t)sm->hist_dist[s] * (dh->dist_lens[s] + dsym_extra(s));
                    fix_bits += (uint64_t)sm->hist_dist[s] * (5 + dsym_extra(s));
                }
                // a forced second literal code has frequency 1 but is never emitted: the estimate is
                // an upper bound by <= 15 bits, the emitted size below is exact
                uint32_t mode = dyn_bits < fix_bits ? 2u : 1u;
                const uint64_t best_bits = dyn_bits < fix_bits ? dyn_bits : fix_bits;
                if ((uint64_t)n + 5 <= (best_bits + 7) / 8) mode = 0;
                sm->mode = mode;
                BitWriter bw{sm->hdr, 0};
                if (mode == 2) {
                    write_dynamic_header(*dh, bw, 1);
                    assign_codes(dh->lit_lens, NLIT, sm->code_lit);
                    assign_codes(dh->dist_lens, NDIST, sm->code_dist);
                } else if (mode == 1) {
                    bw.put(1, 1);
                    bw.put(1, 2);
                    for (int s = 0; s < 288; s++) dh->lit_lens[s] = (uint8_t)fixed_lit_len(s);
                    for (int s = 0; s < 32; s++) dh->dist_lens[s] = 5;
                    assign_codes(dh->lit_lens, 288, sm->code_lit);
                    assign_codes(dh->dist_lens, NDIST, sm->code_dist);
                }
                sm->hdr_bits = (uint32_t)bw.bitpos;
            }
            __syncthreads();
        } else if (t == 0) {
            sm->mode = 0;
        }
        __syncthreads();
        const uint32_t mode = sm->mode;
        PROF(9)

        uint32_t stream_len;
        if (mode != 0) {
            // ---- P9: bits per thread, block scan --------------------------------------------------
            uint32_t bits = 0;
            for (uint32_t r = r0; r < r1; r++) {
                const uint32_t e = s_entry[r];
                if (e == 0xFF) continue;
                const uint32_t pe = ((r << 5) + 32 < n) ? (r << 5) + 32 : n;
                for (uint32_t p = (r << 5) + e; p < pe;) {
                    const Tok tk = token_at(g_match[p], s_data[p]);
                    bits += sm->code_lit[tk.lsym] >> 16;
                    if (tk.is_match) bits += tk.lbits + (sm->code_dist[tk.dsym] >> 16) + tk.dbits;
                    p += tk.step;
                }
            }
            uint32_t tok_bits;
            const uint32_t my_off = sm->hdr_bits + block_excl_scan(bits, sm->warp_tmp, &tok_bits);
            const uint32_t eob = sm->code_lit[EOB];
            const uint32_t total_bits = sm->hdr_bits + tok_bits + (eob >> 16);
            const uint32_t body = (total_bits + 7) >> 3;
            stream_len = hdr_len + body + 4;
            // zero the deflate words (+ trailer spill) so shared boundary words can be OR-ed
            const uint32_t zw = (body + 4 + 3) >> 2;
            for (uint32_t i = t; i < zw; i += T) out32[i] = 0;
            __syncthreads();
            PROF(10)
            // ---- P10: emission ------------------------------------------------------------------------
            if (t == 0) {  // header bits: whole words stored, last partial word OR-ed
                const uint32_t hb = sm->hdr_bits;
                const uint32_t* hwrd = reinterpret_cast<const uint32_t*>(sm->hdr);
                for (uint32_t i = 0; i < (hb >> 5); i++) out32[i] = hwrd[i];
                if (hb & 31) atomicOr(out32 + (hb >> 5), hwrd[hb >> 5]);
            }
            Emitter em;
            em.begin(out32, my_off);
            for (uint32_t r = r0; r < r1; r++) {
                const uint32_t e = s_entry[r];
                if (e == 0xFF) continue;
                const uint32_t pe = ((r << 5) + 32 < n) ? (r << 5) + 32 : n;
                for (uint32_t p = (r << 5) + e; p < pe;) {
                    const Tok tk = token_at(g_match[p], s_data[p]);
                    const uint32_t lc = sm->code_lit[tk.lsym];
                    em.put(lc & 0xffffu, lc >> 16);
                    if (tk.is_match) {
                        if (tk.lbits) em.put(tk.lval, tk.lbits);
                        const uint32_t dc = sm->code_dist[tk.dsym];
                        em.put(dc & 0xffffu, dc >> 16);
                        if (tk.dbits) em.put(tk.dval, tk.dbits);
                    }
                    p += tk.step;
                }
            }
            em.end();
            if (t == T - 1) {  // the last thread's range list ends the stream: EOB follows all tokens
                Emitter ee;
                ee.begin(out32, sm->hdr_bits + tok_bits);
                ee.put(eob & 0xffffu, eob >> 16);
                ee.end();
            }
            __syncthreads();
            if (t == 0) {
                uint8_t* tr = slot + 2 + hdr_len + body;
                const uint32_t ad = (sm->adler_b << 16) | sm->adler_a;
                tr[0] = (uint8_t)(ad >> 24);
                tr[1] = (uint8_t)(ad >> 16);
                tr[2] = (uint8_t)(ad >> 8);
                tr[3] = (uint8_t)ad;
            }
        } else {
            // ---- stored block(s): n <= 32768 -> a single block ---------------------------------------
            uint8_t* o = slot + 2 + hdr_len;
            if (t == 0) {
                o[0] = 1;
                o[1] = (uint8_t)n;
                o[2] = (uint8_t)(n >> 8);
                o[3] = (uint8_t)~n;
                o[4] = (uint8_t)(~n >> 8);
                const uint32_t ad = (sm->adler_b << 16) | sm->adler_a;
                uint8_t* tr = o + 5 + n;
                tr[0] = (uint8_t)(ad >> 24);
                tr[1] = (uint8_t)(ad >> 16);
                tr[2] = (uint8_t)(ad >> 8);
                tr[3] = (uint8_t)ad;
            }
            for (uint32_t p = t; p < n; p += T) o[5 + p] = s_data[p];
            stream_len = hdr_len + 5 + n + 4;
        }
        if (t == 0) {
            uint8_t cmf, flg;
            zlib_header(a.dict_len != 0, cmf, flg);
            slot[2] = cmf;
            slot[3] = flg;
            if (a.dict_len) {
                slot[4] = (uint8_t)(a.dict_adler >> 24);
                slot[5] = (uint8_t)(a.dict_adler >> 16);
                slot[6] = (uint8_t)(a.dict_adler >> 8);
                slot[7] = (uint8_t)a.dict_adler;
            }
            a.sizes[k] = stream_len;
        }
        PROF(11)
        if (t == 0) atomicAdd(&g_prof[15], 1ull);
    }
}

// ---- helpers around the main kernel ---------------------------------------------------------

// Per selected chunk: worst-case slot size and size class list.  class 0: <= NMAX_SMALL,
// class 1: <= NMAX_LARGE, class 2: longer (multi-block path).
__global__ void classify_kernel(uint64_t start0, const uint64_t* __restrict__ cuts, const uint64_t* __restrict__ select,
                                uint64_t m, uint64_t* __restrict__ slot_size, uint32_t* __restrict__ lists,
                                uint32_t* __restrict__ list_n) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint64_t j = select ? select[k] : k;
    const uint64_t s = j ? cuts[j - 1] : start0;
    const uint64_t len = cuts[j] - s;
    const uint64_t bound = len + 5 * (len / 65535 + 1) + 6 + 4 + 2 + 16;
    slot_size[k] = (bound + 15) & ~15ull;
    const int c = len <= NMAX_SMALL ? 0 : (len <= NMAX_LARGE ? 1 : 2);
    const uint32_t pos = atomicAdd(&list_n[c], 1u);
    lists[(uint64_t)c * m + pos] = (uint32_t)k;
}

// Chunks longer than 32 KiB: stored blocks of <= 65535 bytes (valid zlib stream, ratio 1).
__global__ void __launch_bounds__(256)
stored_kernel(DeflArgs a) {
    __shared__ uint32_t s_a[256], s_b[256];
    const uint32_t list_n = *a.list_n;
    for (uint32_t job = blockIdx.x; job < list_n; job += gridDim.x) {
        const uint32_t k = a.list[job];
        const uint64_t j = a.select ? a.select[k] : (uint64_t)k;
        const uint64_t cs = j ? a.cuts[j - 1] : a.start0;
        const uint64_t n = a.cuts[j] - cs;
        const uint8_t* src = a.data + cs;
        uint8_t* slot = a.stage + a.slot_off[k];
        const uint32_t hdr_len = a.dict_len ? 6 : 2;
        uint8_t* o = slot + 2 + hdr_len;
        const uint64_t nblk = n / 65535 + 1;  // the last block may be empty-length only when n % 65535 == 0
        // Adler-32 in segments of 256 bytes per thread-step
        uint64_t sa = 0, sb = 0;
        for (uint64_t p = threadIdx.x; p < n; p += 256) {
            const uint64_t b = src[p];
            sa += b;
            sb = (sb + ((n - p) % 65521u) * b) % 65521u;
            const uint64_t blk = p / 65535;
            o[5 * (blk + 1) + p] = (uint8_t)b;
        }
        s_a[threadIdx.x] = (uint32_t)(sa % 65521u);
        s_b[threadIdx.x] = (uint32_t)sb;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t ta = 1, tb = n % 65521u;
            for (int i = 0; i < 256; i++) {
                ta += s_a[i];
                tb += s_b[i];
            }
            const uint32_t ad = (uint32_t)((tb % 65521u) << 16) | (uint32_t)(ta % 65521u);
            for (uint64_t b = 0; b < nblk; b++) {
                const uint64_t off = b * 65535;
                const uint32_t bl = (uint32_t)(n - off < 65535 ? n - off : 65535);
                uint8_t* h = o + off + 5 * b;
                h[0] = b == nblk - 1 ? 1 : 0;
                h[1] = (uint8_t)bl;
                h[2] = (uint8_t)(bl >> 8);
                h[3] = (uint8_t)~bl;
                h[4] = (uint8_t)(~bl >> 8);
            }
            uint8_t* tr = o + n + 5 * nblk;
            tr[0] = (uint8_t)(ad >> 24);
            tr[1] = (uint8_t)(ad >> 16);
            tr[2] = (uint8_t)(ad >> 8);
            tr[3] = (uint8_t)ad;
            uint8_t cmf, flg;
            zlib_header(a.dict_len != 0, cmf, flg);
            slot[2] = cmf;
            slot[3] = flg;
            if (a.dict_len) {
                slot[4] = (uint8_t)(a.dict_adler >> 24);
                slot[5] = (uint8_t)(a.dict_adler >> 16);
                slot[6] = (uint8_t)(a.dict_adler >> 8);
                slot[7] = (uint8_t)a.dict_adler;
            }
            a.sizes[k] = hdr_len + n + 5 * nblk + 4;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ stage, const uint64_t* __restrict__ slot_off, const uint64_t* __restrict__ offs,
            uint64_t m, uint8_t* __restrict__ out, uint64_t out_cap) {
    for (uint64_t k = blockIdx.x; k < m; k += gridDim.x) {
        const uint8_t* s = stage + slot_off[k] + 2;
        const uint64_t o = offs[k], len = offs[k + 1] - o;
        if (o + len > out_cap) continue;
        uint8_t* d = out + o;
        // head bytes until d is 4-aligned, then words assembled from the (2-aligned) source
        uint64_t head = (4 - ((uintptr_t)d & 3)) & 3;
        if (head > len) head = len;
        if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
        const uint64_t body = (len - head) >> 2;
        const uint8_t* sb = s + head;
        const uint32_t kmis = (uint32_t)((uintptr_t)sb & 3);
        const uint32_t* ws = reinterpret_cast<const uint32_t*>(sb - kmis);
        uint32_t* wd = reinterpret_cast<uint32_t*>(d + head);
        for (uint64_t i = threadIdx.x; i < body; i += 256) {
            uint32_t lo = ws[i], hi = kmis ? ws[i + 1] : 0u;
            wd[i] = __funnelshift_r(lo, hi, kmis * 8);
        }
        const uint64_t done = head + (body << 2);
        if (threadIdx.x < len - done) d[done + threadIdx.x] = s[done + threadIdx.x];
    }
}

uint32_t host_adler32(const uint8_t* d, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; i++) {
        a = (a + d[i]) % 65521u;
        b = (b + a) % 65521u;
    }
    return (b << 16) | a;
}

// Dictionary setup (the analogue of zlib's deflateSetDictionary): a <= 32 KiB one-off, indexed on
// the host and cached until the dictionary bytes change.  Not on the per-chunk path.
struct DictHost {
    uint8_t bytes[DICT_MAX];
    uint32_t len;
    uint32_t adler;
    int valid;
};
int ensure_dict(hmse_ctx* ctx, const uint8_t* d_zdict, uint32_t dict_len, uint32_t* adler, cudaStream_t st) {
    HMSE_SCRATCH(ctx, dev, DictDev*, SLOT_DEFLATE_DICT, sizeof(DictDev));
    DictHost* hd = (DictHost*)ctx->dict_host;
    if (!hd) {
        hd = (DictHost*)calloc(1, sizeof(DictHost));
        if (!hd) HMSE_FAIL(ctx, HMSE_E_NOMEM, "dictionary host copy");
        ctx->dict_host = hd;
    }
    static thread_local uint8_t tmp[DICT_MAX];
    HMSE_CUDA(ctx, cudaMemcpyAsync(tmp, d_zdict, dict_len, cudaMemcpyDeviceToHost, st));
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    if (hd->valid && hd->len == dict_len && memcmp(hd->bytes, tmp, dict_len) == 0) {
        *adler = hd->adler;
        return HMSE_OK;
    }
    hd->valid = 0;
    memcpy(hd->bytes, tmp, dict_len);
    hd->len = dict_len;
    hd->adler = host_adler32(tmp, dict_len);
    DictDev* img = (DictDev*)calloc(1, sizeof(DictDev));
    if (!img) HMSE_FAIL(ctx, HMSE_E_NOMEM, "dictionary index");
    memcpy(img->bytes, tmp, dict_len);
    const uint32_t nh = dict_len >= 4 ? dict_len - 3 : 0;
    uint32_t* cnt = (uint32_t*)calloc(NBUCKET + 1, sizeof(uint32_t));
    auto le32 = [&](uint32_t j) {
        return (uint32_t)tmp[j] | ((uint32_t)tmp[j + 1] << 8) | ((uint32_t)tmp[j + 2] << 16) | ((uint32_t)tmp[j + 3] << 24);
    };
    for (uint32_t j = 0; j < nh; j++) cnt[hash4(le32(j)) + 1]++;
    for (uint32_t h = 0; h < NBUCKET; h++) cnt[h + 1] += cnt[h];
    for (uint32_t h = 0; h <= NBUCKET; h++) img->boff[h] = (uint16_t)cnt[h];
    // descending position inside a bucket: walk positions from the end
    for (uint32_t j = nh; j-- > 0;) {
        const uint32_t v = le32(j), h = hash4(v);
        const uint32_t i = cnt[h]++;
        img->sorted[i] = (uint16_t)j;
        img->first4[i] = v;
    }
    free(cnt);
    cudaError_t e = cudaMemcpyAsync(dev, img, sizeof(DictDev), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(img);
    if (e != cudaSuccess) HMSE_FAIL(ctx, HMSE_E_CUDA, "dictionary upload: %s", cudaGetErrorString(e));
    hd->valid = 1;
    *adler = hd->adler;
    return HMSE_OK;
}

size_t smem_for(uint32_t nmax) {
    return (size_t)nmax + 16 + 2 * (size_t)nmax + (NBUCKET / 2 + 4) * 4 + nmax / 32 + sizeof(Small) + 64;
}

}  // namespace

HMSE_API int hmse_debug_deflate_prof(uint64_t* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_prof, sizeof(g_prof)) != cudaSuccess) return HMSE_E_CUDA;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return HMSE_E_CUDA;
    }
    return HMSE_OK;
}

HMSE_API uint64_t hmse_compress_bound(uint64_t len) { return len + 5 * (len / 65535 + 1) + 6 + 4; }

HMSE_API int hmse_compress(hmse_ctx* ctx, const uint8_t* d_data, uint64_t start0, const uint64_t* d_cuts,
                           const uint64_t* d_select, uint64_t m, const uint8_t* d_zdict, uint32_t dict_len, int level,
                           uint8_t* d_out, uint64_t out_cap, uint64_t* d_offsets, uint64_t* total, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (!total || !d_offsets) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_compress: total/d_offsets is null");
    *total = 0;
    if (dict_len > DICT_MAX) HMSE_FAIL(ctx, HMSE_E_INVAL, "dict_len must be <= 32768");
    if (dict_len && !d_zdict) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_zdict is null");
    if (level < 0 || level > 9) HMSE_FAIL(ctx, HMSE_E_INVAL, "level must be 0..9");
    if (m >= 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "too many chunks in one call");
    if (m == 0) {
        HMSE_CUDA(ctx, cudaMemsetAsync(d_offsets, 0, 8, st));
        return HMSE_OK;
    }
    if (!d_data || !d_cuts) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_compress: null pointer");
    if ((uintptr_t)d_data & 3) HMSE_FAIL(ctx, HMSE_E_INVAL, "d_data must be 4-byte aligned");

    uint32_t dict_adler = 0;
    if (dict_len) {
        int rc = ensure_dict(ctx, d_zdict, dict_len, &dict_adler, st);
        if (rc) return rc;
    }
    // misc: [slot_size m][slot_off m][sizes m+1][lists 3m u32][list_n 4 u32][counters 4 u32][total u64]
    const size_t misc_bytes = (3 * m + 2) * 8 + 3 * m * 4 + 64;
    HMSE_SCRATCH(ctx, misc, uint8_t*, SLOT_DEFLATE_MISC, misc_bytes);
    uint64_t* slot_size = (uint64_t*)misc;
    uint64_t* slot_off = slot_size + m;
    uint64_t* sizes = slot_off + m;  // m + 1
    uint32_t* lists = (uint32_t*)(sizes + m + 1);
    uint32_t* list_n = lists + 3 * m;
    unsigned int* counters = list_n + 4;
    uint64_t* d_tot = (uint64_t*)(counters + 4);  // [stage total, out total]
    d_tot = (uint64_t*)(((uintptr_t)d_tot + 7) & ~(uintptr_t)7);
    HMSE_CUDA(ctx, cudaMemsetAsync(list_n, 0, 64, st));
    KL(ctx);
    classify_kernel<<<(unsigned)div_up64(m, 256), 256, 0, st>>>(start0, d_cuts, d_select, m, slot_size, lists, list_n);
    HMSE_LAUNCH_CHECK(ctx);
    int rc = hmse_exclusive_scan_u64(ctx, slot_size, slot_off, m, d_tot, st);
    if (rc) return rc;
    volatile uint64_t* mail = ctx->pinned;
    HMSE_CUDA(ctx, cudaMemcpyAsync((void*)mail, d_tot, 8, cudaMemcpyDeviceToHost, st));
    HMSE_CUDA(ctx, cudaMemcpyAsync((void*)(mail + 1), list_n, 16, cudaMemcpyDeviceToHost, st));
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t stage_bytes = mail[0];
    const uint32_t* hn = (const uint32_t*)(mail + 1);
    const uint32_t n_small = hn[0], n_large = hn[1], n_huge = hn[2];
    HMSE_SCRATCH(ctx, stage, uint8_t*, SLOT_DEFLATE_STAGE, stage_bytes + 64);

    DeflArgs a;
    a.data = d_data;
    a.start0 = start0;
    a.cuts = d_cuts;
    a.select = d_select;
    a.dict = (const DictDev*)ctx->slot[SLOT_DEFLATE_DICT];
    a.dict_len = dict_len;
    a.dict_adler = dict_adler;
    a.level = level;
    a.stage = stage;
    a.slot_off = slot_off;
    a.sizes = sizes;

    const size_t sm_small = smem_for(NMAX_SMALL), sm_large = smem_for(NMAX_LARGE);
    HMSE_CUDA(ctx, cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_large));
    const uint32_t max_small = (uint32_t)ctx->sm_count * 3, max_large = (uint32_t)ctx->sm_count;
    const uint32_t g_small = n_small < max_small ? n_small : max_small;
    const uint32_t g_large = n_large < max_large ? n_large : max_large;
    HMSE_SCRATCH(ctx, work, uint32_t*, SLOT_DEFLATE_WORK,
                 ((size_t)g_small * NMAX_SMALL + (size_t)g_large * NMAX_LARGE) * 4 + 64);
    HT_BEGIN(ctx, HT_DEFLATE, st);
    if (g_small) {
        a.list = lists;
        a.list_n = list_n;
        a.nmax = NMAX_SMALL;
        a.match = work;
        a.counter = counters;
        KL(ctx);
        deflate_kernel<<<g_small, T_SMALL, sm_small, st>>>(a);
        HMSE_LAUNCH_CHECK(ctx);
    }
    if (g_large) {
        a.list = lists + m;
        a.list_n = list_n + 1;
        a.nmax = NMAX_LARGE;
        a.match = work + (size_t)g_small * NMAX_SMALL;
        a.counter = counters + 1;
        KL(ctx);
        deflate_kernel<<<g_large, T_LARGE, sm_large, st>>>(a);
        HMSE_LAUNCH_CHECK(ctx);
    }
    if (n_huge) {
        a.list = lists + 2 * m;
        a.list_n = list_n + 2;
        KL(ctx);
        stored_kernel<<<n_huge < 1024 ? n_huge : 1024, 256, 0, st>>>(a);
        HMSE_LAUNCH_CHECK(ctx);
    }
    HT_END(ctx, HT_DEFLATE, st);
    HMSE_CUDA(ctx, cudaMemsetAsync(sizes + m, 0, 8, st));
    rc = hmse_exclusive_scan_u64(ctx, sizes, d_offsets, m + 1, d_tot + 1, st);
    if (rc) return rc;
    HMSE_CUDA(ctx, cudaMemcpyAsync((void*)mail, d_tot + 1, 8, cudaMemcpyDeviceToHost, st));
    HMSE_CUDA(ctx, cudaStreamSynchronize(st));
    *total = mail[0];
    if (!d_out || mail[0] > out_cap)
        HMSE_FAIL(ctx, HMSE_E_CAPACITY, "d_out capacity %llu < %llu bytes", (unsigned long long)out_cap,
                  (unsigned long long)mail[0]);
    const uint64_t pg = m < (uint64_t)ctx->sm_count * 16 ? m : (uint64_t)ctx->sm_count * 16;
    HT_BEGIN(ctx, HT_PACK, st);
    KL(ctx);
    pack_kernel<<<(unsigned)pg, 256, 0, st>>>(stage, slot_off, d_offsets, m, d_out, out_cap);
    HMSE_LAUNCH_CHECK(ctx);
    HT_END(ctx, HT_PACK, st);
    return HMSE_OK;
}
