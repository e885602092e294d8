/* C restatement of two oracle loops.  TEST INFRASTRUCTURE - see oracle/__init__.py.
 *
 * hmse_ref_chunk   : FastCDC Algorithm 1 (paper cited at README.md:2753-2755), the same
 *                    byte-at-a-time loop as oracle/cdc.py::next_cut, parameters per
 *                    README.md:289, 2444-2446.
 * hmse_ref_minhash : the spec's minhash_compute (README.md:2578-2597) with the public
 *                    MurmurHash3_x86_32 (the spec's un-vendored murmur3.h, README.md:2573).
 *
 * Built by oracle/build_ref.py into oracle/_build/libhmse_ref.so.  Never linked into or
 * loaded by the product library.
 */
#include <stdint.h>
#include <stddef.h>

static uint64_t next_cut(const uint8_t *d, uint64_t start, uint64_t n, uint32_t mn, uint32_t av,
                         uint32_t mx, uint64_t ms, uint64_t ml, const uint64_t *gear) {
    uint64_t rem = n - start;
    if (rem <= mn) return n;
    uint64_t end = start + (rem < mx ? rem : mx);
    uint64_t normal = start + ((end - start) < av ? (end - start) : av);
    uint64_t fp = 0, i = start + mn;
    for (; i < normal; i++) {
        fp = (fp << 1) + gear[d[i]];
        if (!(fp & ms)) return i;
    }
    for (; i < end; i++) {
        fp = (fp << 1) + gear[d[i]];
        if (!(fp & ml)) return i;
    }
    return end;
}

/* Walks chunk starts s with entry <= s < n_own over d[0:n); writes each start's cut.
 * eof != 0: d ends the stream.  Returns the number of cuts, or -1 if cap is too small. */
int64_t hmse_ref_chunk(const uint8_t *d, uint64_t n, uint64_t entry, uint64_t n_own, int eof,
                       uint32_t mn, uint32_t av, uint32_t mx, uint64_t ms, uint64_t ml,
                       const uint64_t *gear, uint64_t *cuts, uint64_t cap) {
    uint64_t k = 0, s = entry;
    if (eof) n_own = n;
    while (s < n_own) {
        s = next_cut(d, s, n, mn, av, mx, ms, ml, gear);
        if (k >= cap) return -1;
        cuts[k++] = s;
    }
    return (int64_t)k;
}

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* MurmurHash3_x86_32 specialised to a 4-byte key (one body block, no tail). */
static inline uint32_t murmur3_4(uint32_t k, uint32_t seed) {
    k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u;
    uint32_t h = seed ^ k;
    h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
    h ^= 4u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

/* sig[c][p] = min over pos in [start_c, cut_c - 3) of murmur3(le32(d+pos), seeds[p]);
 * chunks shorter than 4 bytes keep 0xFFFFFFFF. */
void hmse_ref_minhash(const uint8_t *d, uint64_t start0, const uint64_t *cuts, uint64_t n_chunks,
                      const uint32_t *seeds, uint32_t n_perm, uint32_t *sig) {
    uint64_t s = start0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        uint64_t e = cuts[c];
        uint32_t *out = sig + c * (uint64_t)n_perm;
        for (uint32_t p = 0; p < n_perm; p++) out[p] = 0xFFFFFFFFu;
        for (uint64_t pos = s; pos + 4 <= e; pos++) {
            uint32_t k = (uint32_t)d[pos] | ((uint32_t)d[pos + 1] << 8) | ((uint32_t)d[pos + 2] << 16) |
                         ((uint32_t)d[pos + 3] << 24);
            k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u;
            for (uint32_t p = 0; p < n_perm; p++) {
                uint32_t h = seeds[p] ^ k;
                h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
                h ^= 4u;
                h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
                if (h < out[p]) out[p] = h;
            }
        }
        s = e;
    }
}

uint32_t hmse_ref_murmur3_4(uint32_t k, uint32_t seed) { return murmur3_4(k, seed); }
