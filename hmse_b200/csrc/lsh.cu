// lsh.cu - LSH bucketing (spec: README.md:1375-1383, 2229-2245): every (band, key, chunk id)
// triple sorted by (band, key, id); equal (band, key) runs are the candidate buckets.
//
// The band is the major key and ids start ascending, so the job is `bands` independent STABLE
// sorts of n 64-bit keys: an LSD radix sort, 8 passes of 8 bits, all bands in one launch
// (blockIdx.y = band).  Per pass: tile histograms -> digit-major scan -> stable scatter where
// the in-tile rank comes from warp match_any groups (no atomics, deterministic).
// HBM-bound: 8 passes x (12 B read + 12 B written) per triple.
#include "ctx.cuh"

namespace {

constexpr int LT = 256;            // threads
constexpr int LI = 8;              // items per thread
constexpr int LTILE = LT * LI;     // 2048
constexpr int LW = LT / 32;

struct SortArgs {
    const uint64_t* key_in;   // pass 0: keys[n][bands] (strided); else [band][n]
    const uint32_t* id_in;    // null in pass 0
    uint64_t* key_out;
    uint32_t* id_out;         // passes 0..6
    uint64_t* id_out64;       // last pass: id_base + id
    uint32_t* band_out;       // last pass
    uint32_t* hist;           // [band][digit][tile]
    uint64_t n, id_base;
    uint32_t bands, n_tiles, shift, first, last;
};

__device__ __forceinline__ void load_item(const SortArgs& a, uint32_t band, uint64_t i, uint64_t& k, uint32_t& id) {
    if (a.first) {
        k = a.key_in[i * a.bands + band];
        id = (uint32_t)i;
    } else {
        k = a.key_in[(uint64_t)band * a.n + i];
        id = a.id_in[(uint64_t)band * a.n + i];
    }
}

__global__ void __launch_bounds__(LT) lsh_hist_kernel(SortArgs a) {
    __shared__ uint32_t h[256];
    const uint32_t band = blockIdx.y, tile = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)tile * LTILE;
#pragma unroll
    for (int it = 0; it < LI; it++) {
        const uint64_t i = base + (uint64_t)it * LT + threadIdx.x;
        if (i < a.n) {
            const uint64_t k = a.first ? a.key_in[i * a.bands + band] : a.key_in[(uint64_t)band * a.n + i];
            atomicAdd(&h[(k >> a.shift) & 255], 1u);
        }
    }
    __syncthreads();
    a.hist[((uint64_t)band * 256 + threadIdx.x) * a.n_tiles + tile] = h[threadIdx.x];
}

// One block per band: exclusive scan of hist[band][digit][tile] in digit-major order, in place.
__global__ void __launch_bounds__(LT) lsh_scan_kernel(uint32_t* __restrict__ hist, uint32_t n_tiles) {
    __shared__ uint32_t wsum[LW];
    __shared__ uint32_t carry_s;
    uint32_t* p = hist + (uint64_t)blockIdx.x * 256 * n_tiles;
    const uint64_t total = (uint64_t)256 * n_tiles;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (uint64_t b0 = 0; b0 < total; b0 += LT * 4) {
        const uint64_t i0 = b0 + (uint64_t)threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v[q] = i0 + q < total ? p[i0 + q] : 0;
            s += v[q];
        }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        __syncthreads();
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t x = lane < LW ? wsum[lane] : 0, xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, xi, o);
                if (lane >= (unsigned)o) xi += t;
            }
            if (lane < LW) wsum[lane] = xi - x;
            if (lane == 31) carry_s = xi;
        }
        __syncthreads();
        uint32_t run = carry + wsum[w] + inc - s;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (i0 + q < total) p[i0 + q] = run;
            run += v[q];
        }
        carry += carry_s;
    }
}

__global__ void __launch_bounds__(LT) lsh_scatter_kernel(SortArgs a) {
    __shared__ uint32_t cnt[LW][256];   // per-warp digit counts -> exclusive warp bases
    __shared__ uint32_t gofs[256];      // global offset of this tile's first element per digit
    const uint32_t band = blockIdx.y, tile = blockIdx.x;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < LW * 256; i += LT) (&cnt[0][0])[i] = 0;
    gofs[threadIdx.x] = a.hist[((uint64_t)band * 256 + threadIdx.x) * a.n_tiles + tile];
    __syncthreads();
    // element order inside the tile: warp-major, then item, then lane (stable w.r.t. input index)
    const uint64_t wbase = (uint64_t)tile * LTILE + (uint64_t)w * (32 * LI);
    uint64_t k[LI];
    uint32_t id[LI], rank[LI];
    bool ok[LI];
#pragma unroll
    for (int it = 0; it < LI; it++) {
        const uint64_t i = wbase + (uint64_t)it * 32 + lane;
        ok[it] = i < a.n;
        k[it] = 0;
        id[it] = 0;
        if (ok[it]) load_item(a, band, i, k[it], id[it]);
        const uint32_t d = ok[it] ? (uint32_t)((k[it] >> a.shift) & 255) : 256u;  // 256 = padding group
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & ((1u << lane) - 1));
        uint32_t old = 0;
        if (ok[it]) old = cnt[w][d];
        __syncwarp();
        if (ok[it] && before == 0) cnt[w][d] = old + __popc(peers);
        __syncwarp();
        rank[it] = old + before;
    }
    __syncthreads();
    {   // exclusive prefix over warps for digit = threadIdx.x
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < LW; ww++) {
            uint32_t c = cnt[ww][threadIdx.x];
            cnt[ww][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < LI; it++) {
        if (!ok[it]) continue;
        const uint32_t d = (uint32_t)((k[it] >> a.shift) & 255);
        const uint64_t pos = (uint64_t)band * a.n + gofs[d] + cnt[w][d] + rank[it];
        a.key_out[pos] = k[it];
        if (a.last) {
            a.id_out64[pos] = a.id_base + id[it];
            a.band_out[pos] = band;
        } else {
            a.id_out[pos] = id[it];
        }
    }
}

}  // namespace

HMSE_API int hmse_lsh_buckets(hmse_ctx* ctx, const uint64_t* d_keys, uint64_t n, uint32_t bands, uint64_t id_base,
                              uint32_t* d_band, uint64_t* d_key, uint64_t* d_id, void* stream) {
    if (!ctx) return HMSE_E_INVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (bands == 0 || bands > 65535) HMSE_FAIL(ctx, HMSE_E_INVAL, "bands must be in 1..65535");
    if (n == 0) return HMSE_OK;
    if (!d_keys || !d_band || !d_key || !d_id) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_buckets: null pointer");
    if (n > 0xFFFFFFFFull) HMSE_FAIL(ctx, HMSE_E_INVAL, "hmse_lsh_buckets: n exceeds 2^32");
    const uint64_t total = n * bands;
    const uint32_t n_tiles = (uint32_t)div_up64(n, LTILE);
    // ping-pong buffers: keys A, keys B, ids A, ids B
    HMSE_SCRATCH(ctx, buf, uint8_t*, SLOT_LSH_SORT, total * (8 + 8 + 4 + 4) + 64);
    uint64_t* kA = (uint64_t*)buf;
    uint64_t* kB = kA + total;
    uint32_t* iA = (uint32_t*)(kB + total);
    uint32_t* iB = iA + total;
    HMSE_SCRATCH(ctx, hist, uint32_t*, SLOT_LSH_MISC, (size_t)bands * 256 * n_tiles * 4);
    const dim3 grid(n_tiles, bands);
    HT_BEGIN(ctx, HT_LSH, st);
    for (uint32_t pass = 0; pass < 8; pass++) {
        SortArgs a;
        a.first = pass == 0;
        a.last = pass == 7;
        a.key_in = pass == 0 ? d_keys : ((pass & 1) ? kA : kB);
        a.id_in = pass == 0 ? nullptr : ((pass & 1) ? iA : iB);
        a.key_out = a.last ? d_key : ((pass & 1) ? kB : kA);
        a.id_out = (pass & 1) ? iB : iA;
        a.id_out64 = d_id;
        a.band_out = d_band;
        a.hist = hist;
        a.n = n;
        a.id_base = id_base;
        a.bands = bands;
        a.n_tiles = n_tiles;
        a.shift = pass * 8;
        KL(ctx);
        lsh_hist_kernel<<<grid, LT, 0, st>>>(a);
        KL(ctx);
        lsh_scan_kernel<<<bands, LT, 0, st>>>(hist, n_tiles);
        KL(ctx);
        lsh_scatter_kernel<<<grid, LT, 0, st>>>(a);
        HMSE_LAUNCH_CHECK(ctx);
    }
    HT_END(ctx, HT_LSH, st);
    return HMSE_OK;
}
