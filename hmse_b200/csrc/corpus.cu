// corpus.cu - device twin of the procedural templated-wiki corpus (bench/test INPUT generator;
// not part of the reference path).  Every byte is a pure function of (seed, article, token)
// built from 32-bit integer hashes, so 10-100 GB streams are synthesised in HBM and any prefix
// can be regenerated on the CPU by oracle/corpus.py, which this file must match bit for bit
// (tests/test_gpu_corpus.py).  Corpus shape per the spec: templated infobox/cite/category text
// with exact- and near-duplicate articles (README.md:1176-1178, 2123-2127), seed 42
// (VALIDATION_METHODS.md:119-120).
#include "ctx.cuh"

namespace {

constexpr uint32_t NI = 12, NBODY = 12, NPUNCT = 4;
constexpr uint32_t ID_BODY0 = NI, ID_PUNCT0 = NI + NBODY, ID_TERM = ID_PUNCT0 + NPUNCT, ID_WORD0 = ID_TERM + 1;
constexpr uint32_t N_WORDS = 4096;
constexpr uint32_t K_TOKEN = 0x5BD1E995, K_CLASS = 0x000A11CE, K_PICK = 0x00000D0B, K_EDIT = 0x0000ED17,
                   K_NTOK = 0x0000070C, K_EPOS = 0x00001234, K_EWORD = 0x00004321;

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t H(uint32_t seed, uint32_t a, uint32_t b) {
    return mix32(mix32(seed ^ (a * 0x9E3779B1u)) + b * 0x85EBCA77u);
}
__host__ __device__ __forceinline__ uint32_t zipf_word(uint32_t v) {
    const uint32_t a = v & 0xFF, b = (v >> 8) & 0xFF, c = (v >> 16) & 0x3FF;
    return (((a * b) >> 8) * c) >> 6;
}

struct Meta {
    uint32_t c, e, nt, ne;
    uint32_t epos[4], eword[4];
};

__device__ __forceinline__ Meta article_meta(const hmse_corpus_cfg& cfg, uint64_t a64) {
    Meta m;
    const uint32_t a = (uint32_t)a64;
    const uint32_t thr = cfg.dup_thr + cfg.near_thr;
    m.c = a;
    m.e = 0;
    if (a != 0) {
        const uint32_t cls = H(cfg.seed, a, K_CLASS) & 1023;
        if (cls < thr) {
            uint32_t c = 0;
            const uint32_t tries = cfg.pick_tries ? cfg.pick_tries : 4u;
            for (uint32_t k = 0; k < tries; k++) {
                c = (uint32_t)(((uint64_t)H(cfg.seed, a, K_PICK + k) * a) >> 32);
                if (c == 0 || (H(cfg.seed, c, K_CLASS) & 1023) >= thr) break;
            }
            m.c = c;
            if (cls >= cfg.dup_thr) m.e = H(cfg.seed, a, K_EDIT) | 1u;
        }
    }
    m.nt = 800 + (H(cfg.seed, m.c, K_NTOK) & 16383);
    m.ne = m.e ? 1 + (m.e & 3) : 0;
    for (uint32_t k = 0; k < 4; k++) {
        m.epos[k] = k < m.ne ? H(m.e, k, K_EPOS) % m.nt : 0xFFFFFFFFu;
        m.eword[k] = ID_WORD0 + zipf_word(H(m.e, k, K_EWORD) >> 6);
    }
    return m;
}

__device__ __forceinline__ uint32_t token_id(const Meta& m, uint32_t base, uint32_t t) {
    const uint32_t u = mix32(base + t * 0x85EBCA77u);
    const uint32_t v = u >> 6, sel = u & 63;
    const uint32_t word = ID_WORD0 + zipf_word(v);
    uint32_t id = sel == 0 ? ID_BODY0 + v % NBODY : (sel <= 6 ? ID_PUNCT0 + (v & 3) : word);
    if (t < 2 * NI) id = (t & 1) == 0 ? (t >> 1) : word;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (m.epos[k] == t) id = m.eword[k];  // later edits override earlier ones
    if (t == m.nt - 1) id = ID_TERM;
    return id;
}

__global__ void __launch_bounds__(128)
corpus_lengths_kernel(hmse_corpus_cfg cfg, const uint32_t* __restrict__ lex_off, uint64_t first, uint64_t n_art,
                      uint32_t* __restrict__ art_len) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t w = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= n_art) return;
    const Meta m = article_meta(cfg, first + w);
    const uint32_t base = mix32((cfg.seed ^ K_TOKEN) ^ (m.c * 0x9E3779B1u));
    uint32_t sum = 0;
    for (uint32_t t = lane; t < m.nt; t += 32) {
        const uint32_t id = token_id(m, base, t);
        sum += __ldg(lex_off + id + 1) - __ldg(lex_off + id);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) art_len[w] = sum;
}

constexpr int RT = 256;

// One CTA per article: tiles of 256 tokens, block scan of token lengths, byte copies.
__global__ void __launch_bounds__(RT)
corpus_render_kernel(hmse_corpus_cfg cfg, const uint8_t* __restrict__ lex_blob, const uint32_t* __restrict__ lex_off,
                     uint64_t first, uint64_t n_art, const uint64_t* __restrict__ art_off, uint64_t byte_off,
                     uint64_t n, uint8_t* __restrict__ out) {
    __shared__ uint32_t wsum[RT / 32];
    __shared__ uint32_t tile_total;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t ai = blockIdx.x; ai < n_art; ai += gridDim.x) {
        const uint64_t a0 = art_off[ai], a1 = art_off[ai + 1];
        if (a1 <= byte_off || a0 >= byte_off + n) continue;
        const Meta m = article_meta(cfg, first + ai);
        const uint32_t base = mix32((cfg.seed ^ K_TOKEN) ^ (m.c * 0x9E3779B1u));
        uint64_t pos = a0;  // stream offset of the current tile
        for (uint32_t t0 = 0; t0 < m.nt; t0 += RT) {
            const uint32_t t = t0 + threadIdx.x;
            uint32_t id = 0, len = 0, src = 0;
            if (t < m.nt) {
                id = token_id(m, base, t);
                src = __ldg(lex_off + id);
                len = __ldg(lex_off + id + 1) - src;
            }
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += v;
            }
            __syncthreads();
            if (lane == 31) wsum[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                uint32_t s = lane < RT / 32 ? wsum[lane] : 0, si = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, si, o);
                    if (lane >= (unsigned)o) si += v;
                }
                if (lane < RT / 32) wsum[lane] = si - s;
                if (lane == 31) tile_total = si;
            }
            __syncthreads();
            const uint64_t p = pos + wsum[wid] + inc - len;
            for (uint32_t i = 0; i < len; i++) {
                const uint64_t q = p + i;
                if (q >= byte_off && q < byte_off + n) out[q - byte_off] = __ldg(lex_blob + src + i);
            }
            pos += tile_total;
        }
        __syncthreads();
    }
}

}  // namespace

HMSE_API int hmse_corpus_lengths(hmse_ctx* ctx, const hmse_corpus_cfg* cfg, const uint32_t* d_lex_off,
                                 uint64_t first_article, uint64_t n_articles, uint32_t* d_art_len, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (!cfg || !d_lex_off || !d_art_len) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_corpus_lengths: null pointer");
    if (cfg->n_lex != ID_WORD0 + N_WORDS) HMSE_FAIL(ctx, HMSE_E_INVAL, "lexicon must have %u entries", ID_WORD0 + N_WORDS);
    if (first_article + n_articles > 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "article index exceeds 2^32");
    if (n_articles == 0) return HMSE_OK;
    KL(ctx);
    corpus_lengths_kernel<<<(unsigned)div_up64(n_articles, 4), 128, 0, (cudaStream_t)stream>>>(
        *cfg, d_lex_off, first_article, n_articles, d_art_len);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}

HMSE_API int hmse_corpus_render(hmse_ctx* ctx, const hmse_corpus_cfg* cfg, const uint8_t* d_lex_blob,
                                const uint32_t* d_lex_off, uint64_t first_article, uint64_t n_articles,
                                const uint64_t* d_art_off, uint64_t byte_off, uint64_t n, uint8_t* d_out, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    if (!cfg || !d_lex_blob || !d_lex_off || !d_art_off || (!d_out && n))
        HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_corpus_render: null pointer");
    if (cfg->n_lex != ID_WORD0 + N_WORDS) HMSE_FAIL(ctx, HMSE_E_INVAL, "lexicon must have %u entries", ID_WORD0 + N_WORDS);
    if (n_articles == 0 || n == 0) return HMSE_OK;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    KL(ctx);
    corpus_render_kernel<<<(unsigned)(n_articles < cap ? n_articles : cap), RT, 0, (cudaStream_t)stream>>>(
        *cfg, d_lex_blob, d_lex_off, first_article, n_articles, d_art_off, byte_off, n, d_out);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}
