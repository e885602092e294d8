// deflate_core.h - the sequential pieces of the DEFLATE encoder (RFC 1951 symbol maps, Huffman
// code construction with zlib-compatible length limiting, dynamic-block header) as plain
// host+device functions.  The CUDA kernel (deflate.cu) calls them from one thread per chunk;
// tests/model/deflate_model.cpp compiles the same header with g++ so the bit-level logic is
// checked against stock zlib on the CPU before it ever runs on a GPU.
//
// Replaces the spec's miniz calls (README.md:2374, 2378) per SURVEY.md §0.2 C5.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DFL_HD __host__ __device__ __forceinline__
#else
#define DFL_HD inline
#endif

namespace dfl {

constexpr int MIN_MATCH = 3;
constexpr int MAX_MATCH = 258;
constexpr int WSIZE = 32768;
constexpr int NLIT = 286;   // literal/length symbols 0..285
constexpr int NDIST = 30;
constexpr int NCL = 19;
constexpr int EOB = 256;
constexpr int HASH_BITS = 13;
constexpr int NBUCKET = 1 << HASH_BITS;
// zlib level-6 tuning (deflate.c configuration_table[6])
constexpr int GOOD_LENGTH = 8, MAX_LAZY = 16, NICE_LENGTH = 128, MAX_CHAIN = 128;
constexpr int TOO_FAR = 4096;

DFL_HD int ilog2(uint32_t v) {
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// length 3..258 -> symbol 257..285, extra bit count, extra value
DFL_HD void len_sym(uint32_t len, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {
    uint32_t l = len - 3;
    if (l < 8) {
        sym = 257 + l;
        ebits = 0;
        eval = 0;
    } else if (len == 258) {
        sym = 285;
        ebits = 0;
        eval = 0;
    } else {
        uint32_t eb = (uint32_t)ilog2(l) - 2;
        sym = 257 + 4 * eb + 4 + ((l >> eb) & 3);
        ebits = eb;
        eval = l & ((1u << eb) - 1);
    }
}

// distance 1..32768 -> symbol 0..29, extra bit count, extra value
DFL_HD void dist_sym(uint32_t dist, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {
    uint32_t d = dist - 1;
    if (d < 4) {
        sym = d;
        ebits = 0;
        eval = 0;
    } else {
        uint32_t eb = (uint32_t)ilog2(d) - 1;
        sym = 2 * eb + 2 + ((d >> eb) & 1);
        ebits = eb;
        eval = d & ((1u << eb) - 1);
    }
}

DFL_HD uint32_t hash4(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - HASH_BITS); }

DFL_HD uint32_t bitrev(uint32_t code, int len) {
#ifdef __CUDA_ARCH__
    return __brev(code) >> (32 - len);
#endif
    uint32_t r = 0;
    for (int i = 0; i < len; i++) {
        r = (r << 1) | (code & 1);
        code >>= 1;
    }
    return r;
}

// Sequential LSB-first bit writer (RFC 1951 packing) into a byte buffer; starts byte aligned.
// finish() flushes the last partial byte; bitpos is the number of bits written.
struct BitWriter {
    uint8_t* buf;
    uint64_t bitpos;
    uint64_t acc = 0;
    int nacc = 0;
    DFL_HD void put(uint32_t value, int nbits) {
        acc |= (uint64_t)value << nacc;
        nacc += nbits;
        while (nacc >= 8) {
            buf[bitpos >> 3] = (uint8_t)acc;
            acc >>= 8;
            nacc -= 8;
            bitpos += 8;
        }
    }
    DFL_HD void finish() {
        if (nacc) {
            buf[bitpos >> 3] = (uint8_t)acc;
            bitpos += (uint64_t)nacc;
            nacc = 0;
            acc = 0;
        }
    }
};

// Scratch for one Huffman construction over at most CAP symbols.
template <int CAP>
struct HuffWorkT {
    uint32_t w[2 * CAP];      // node weights: leaves (sorted ascending) then internal nodes
    uint16_t parent[2 * CAP];
    uint16_t order[CAP];      // leaf i (sorted position) -> symbol
    uint16_t bl_count[18];
};
using HuffWork = HuffWorkT<288>;
using HuffWorkSmall = HuffWorkT<32>;

// Length-limited Huffman code lengths from k >= 2 leaves already sorted ascending by
// (freq, symbol) in hw.w[0..k) / hw.order[0..k).  lens[0..n) receives 0 for unused symbols.
// Depth overflow is repaired with the bl_count move zlib's gen_bitlen uses, so the code is
// always complete (inflate rejects incomplete sets).
// Two-queue Huffman merge over k sorted leaves: leaves 0..k-1, internal nodes k..2k-2 (created in
// ascending weight order, so parents always have larger indices); root = 2k-2.
template <class HW>
DFL_HD void huff_merge(HW& hw, int k) {
    int leaf = 0, inode = k, next = k;
    for (int i = 0; i < k - 1; i++) {
        int a, b;
        if (leaf < k && (inode >= next || hw.w[leaf] <= hw.w[inode])) a = leaf++; else a = inode++;
        if (leaf < k && (inode >= next || hw.w[leaf] <= hw.w[inode])) b = leaf++; else b = inode++;
        hw.w[next] = hw.w[a] + hw.w[b];
        hw.parent[a] = (uint16_t)next;
        hw.parent[b] = (uint16_t)next;
        next++;
    }
}

// zlib gen_bitlen's overflow repair on a per-length leaf histogram whose Kraft sum (in units of
// 2^-maxbits) is `kraft`: afterwards the histogram describes a complete prefix code.
DFL_HD void huff_fix_overflow(uint16_t* bl_count, int maxbits, uint64_t kraft) {
    while (kraft > (1ull << maxbits)) {
        int bits = maxbits - 1;
        while (bl_count[bits] == 0) bits--;
        bl_count[bits]--;
        bl_count[bits + 1] += 2;
        bl_count[maxbits]--;
        kraft--;
    }
}

template <class HW>
DFL_HD void build_lengths_sorted(HW& hw, int k, int maxbits, uint8_t* lens, int n) {
    for (int s = 0; s < n; s++) lens[s] = 0;
    huff_merge(hw, k);
    const int root = 2 * k - 2;
    // depths: w[] of internal nodes becomes depth
    hw.w[root] = 0;
    for (int i = root - 1; i >= k; i--) hw.w[i] = hw.w[hw.parent[i]] + 1;
    for (int b = 0; b <= maxbits; b++) hw.bl_count[b] = 0;
    uint64_t kraft = 0;  // in units of 2^-maxbits
    for (int i = 0; i < k; i++) {
        uint32_t d = hw.w[hw.parent[i]] + 1;
        if (d > (uint32_t)maxbits) d = (uint32_t)maxbits;
        hw.bl_count[d]++;
        kraft += 1ull << (maxbits - d);
    }
    huff_fix_overflow(hw.bl_count, maxbits, kraft);
    // rarest symbols take the longest codes
    int i = 0;
    for (int bits = maxbits; bits >= 1; bits--)
        for (int c = hw.bl_count[bits]; c > 0; c--) lens[hw.order[i++]] = (uint8_t)bits;
}

// Same from raw frequencies: forces at least two codes (zlib trees.c build_tree does the same)
// and insertion-sorts the used symbols by (freq, symbol).  freq[] may be modified.
template <class HW>
DFL_HD void build_lengths(uint32_t* freq, int n, int maxbits, uint8_t* lens, HW& hw) {
    int k = 0;
    for (int s = 0; s < n; s++)
        if (freq[s]) k++;
    if (k < 2) {
        if (k == 0) {
            freq[0] = 1;
            freq[1] = 1;
        } else if (freq[0]) {
            freq[1] = 1;
        } else {
            freq[0] = 1;
        }
        k = 2;
    }
    int m = 0;
    for (int s = 0; s < n; s++) {
        if (!freq[s]) continue;
        int i = m++;
        while (i > 0 && hw.w[i - 1] > freq[s]) {
            hw.w[i] = hw.w[i - 1];
            hw.order[i] = hw.order[i - 1];
            i--;
        }
        hw.w[i] = freq[s];
        hw.order[i] = (uint16_t)s;
    }
    build_lengths_sorted(hw, k, maxbits, lens, n);
}

// Canonical codes (RFC 1951 3.2.2), bit-reversed for LSB-first emission; codes[s] = len<<16 | code.
DFL_HD void assign_codes(const uint8_t* lens, int n, uint32_t* codes) {
    uint16_t bl[16], nc[16];
    for (int b = 0; b < 16; b++) bl[b] = 0;
    for (int s = 0; s < n; s++) bl[lens[s]]++;
    bl[0] = 0;
    uint32_t code = 0;
    nc[0] = 0;
    for (int b = 1; b < 16; b++) {
        code = (code + bl[b - 1]) << 1;
        nc[b] = (uint16_t)code;
    }
    for (int s = 0; s < n; s++) {
        int l = lens[s];
        codes[s] = l ? ((uint32_t)l << 16) | bitrev(nc[l]++, l) : 0;
    }
}

// Order in which code-length-code lengths are transmitted (RFC 1951 3.2.7).
DFL_HD int cl_order(int i) {
    const uint8_t order[NCL] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    return order[i];
}

// Run-length coding of one tree's code lengths into code-length symbols (zlib scan_tree /
// send_tree).  Each output entry = sym | extra_value << 8.  Returns the entry count.
DFL_HD int rle_lengths(const uint8_t* lens, int n, uint16_t* out) {
    int m = 0;
    int i = 0;
    while (i < n) {
        int cur = lens[i];
        int run = 1;
        while (i + run < n && lens[i + run] == cur) run++;
        i += run;
        if (cur == 0) {
            while (run >= 11) {
                int r = run > 138 ? 138 : run;
                out[m++] = (uint16_t)(18 | ((r - 11) << 8));
                run -= r;
            }
            if (run >= 3) {
                out[m++] = (uint16_t)(17 | ((run - 3) << 8));
                run = 0;
            }
            while (run-- > 0) out[m++] = 0;
        } else {
            out[m++] = (uint16_t)cur;
            run--;
            while (run >= 3) {
                int r = run > 6 ? 6 : run;
                out[m++] = (uint16_t)(16 | ((r - 3) << 8));
                run -= r;
            }
            while (run-- > 0) out[m++] = (uint16_t)cur;
        }
    }
    return m;
}

struct DynHeader {
    uint8_t lit_lens[288];
    uint8_t dist_lens[32];
    uint8_t cl_lens[NCL];
    uint32_t cl_codes[NCL];
    uint16_t rle[288 + 32];
    int n_rle, nlit, ndist, ncl;
    uint32_t bits;  // header bit count: 3 + 14 + 3*ncl + coded lengths
};

// Header plan from finished lit/dist code lengths: trims trailing zeros, run-length codes the
// lengths, builds the code-length code.  Returns the header size in bits incl. BFINAL/BTYPE.
template <class HW>
DFL_HD uint32_t plan_header_from_lengths(DynHeader& h, HW& hw) {
    h.nlit = NLIT;
    while (h.nlit > 257 && h.lit_lens[h.nlit - 1] == 0) h.nlit--;
    h.ndist = NDIST;
    while (h.ndist > 1 && h.dist_lens[h.ndist - 1] == 0) h.ndist--;
    int m = rle_lengths(h.lit_lens, h.nlit, h.rle);
    m += rle_lengths(h.dist_lens, h.ndist, h.rle + m);
    h.n_rle = m;
    uint32_t cl_freq[NCL];
    for (int i = 0; i < NCL; i++) cl_freq[i] = 0;
    for (int i = 0; i < m; i++) cl_freq[h.rle[i] & 0xff]++;
    build_lengths(cl_freq, NCL, 7, h.cl_lens, hw);
    assign_codes(h.cl_lens, NCL, h.cl_codes);
    h.ncl = NCL;
    while (h.ncl > 4 && h.cl_lens[cl_order(h.ncl - 1)] == 0) h.ncl--;
    uint32_t bits = 3 + 5 + 5 + 4 + 3 * (uint32_t)h.ncl;
    for (int i = 0; i < m; i++) {
        int s = h.rle[i] & 0xff;
        bits += h.cl_lens[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
    }
    h.bits = bits;
    return bits;
}

// lit_freq / dist_freq may be modified (two-code forcing).
DFL_HD uint32_t plan_dynamic_header(uint32_t* lit_freq, uint32_t* dist_freq, DynHeader& h, HuffWork& hw) {
    build_lengths(lit_freq, NLIT, 15, h.lit_lens, hw);
    build_lengths(dist_freq, NDIST, 15, h.dist_lens, hw);
    return plan_header_from_lengths(h, hw);
}

DFL_HD void write_dynamic_header(const DynHeader& h, BitWriter& bw, int bfinal) {
    bw.put((uint32_t)bfinal, 1);
    bw.put(2, 2);
    bw.put((uint32_t)(h.nlit - 257), 5);
    bw.put((uint32_t)(h.ndist - 1), 5);
    bw.put((uint32_t)(h.ncl - 4), 4);
    for (int i = 0; i < h.ncl; i++) bw.put(h.cl_lens[cl_order(i)], 3);
    for (int i = 0; i < h.n_rle; i++) {
        int s = h.rle[i] & 0xff, ev = h.rle[i] >> 8;
        bw.put(h.cl_codes[s] & 0xffff, (int)(h.cl_codes[s] >> 16));
        if (s == 16) bw.put((uint32_t)ev, 2);
        else if (s == 17) bw.put((uint32_t)ev, 3);
        else if (s == 18) bw.put((uint32_t)ev, 7);
    }
}

// Fixed-Huffman codes (RFC 1951 3.2.6), bit-reversed, packed len<<16 | code.
DFL_HD uint32_t fixed_lit_code(int s) {
    if (s < 144) return (8u << 16) | bitrev(0x30u + (uint32_t)s, 8);
    if (s < 256) return (9u << 16) | bitrev(0x190u + (uint32_t)(s - 144), 9);
    if (s < 280) return (7u << 16) | bitrev((uint32_t)(s - 256), 7);
    return (8u << 16) | bitrev(0xC0u + (uint32_t)(s - 280), 8);
}
DFL_HD uint32_t fixed_dist_code(int s) { return (5u << 16) | bitrev((uint32_t)s, 5); }

// Fixed-Huffman code lengths (RFC 1951 3.2.6).
DFL_HD int fixed_lit_len(int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }

// Extra bits carried by a length / distance symbol.
DFL_HD int lsym_extra(int s) { return (s < 265 || s == 285) ? 0 : (s - 261) >> 2; }
DFL_HD int dsym_extra(int s) { return s < 4 ? 0 : (s - 2) >> 1; }

// zlib stream header bytes (RFC 1950): CMF 0x78, FLEVEL 2, FDICT as given.
DFL_HD void zlib_header(int has_dict, uint8_t& cmf, uint8_t& flg) {
    cmf = 0x78;
    uint32_t f = 0x80 | (has_dict ? 0x20 : 0);
    f += (31 - ((0x78u * 256 + f) % 31)) % 31;
    flg = (uint8_t)f;
}

}  // namespace dfl
