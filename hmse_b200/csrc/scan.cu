// scan.cu - generic device-wide exclusive prefix sum over u64 (three small kernels: tile
// reduce, single-block scan of the tile sums, tile scan + offset).  Used for cut-list
// compaction, compressed-stream offsets and partition offsets; never on a per-byte path.
#include "ctx.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// Inclusive scan across the block; returns this thread's inclusive value, *total = block sum.
__device__ __forceinline__ uint64_t block_incl_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t wsum[SCAN_THREADS / 32];
    __shared__ uint64_t wtot;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = warp_incl_scan(v);
    __syncthreads();  // protect wsum reuse across calls
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint64_t s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        uint64_t si = warp_incl_scan(s);
        if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) wtot = si;
    }
    __syncthreads();
    *total = wtot;
    return inc + wsum[w];
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint64_t* __restrict__ in, uint64_t n,
                                                               uint64_t* __restrict__ sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) v += in[base + i];
    uint64_t tot;
    block_incl_scan(v, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// One block scans all tile sums in place (exclusive) and writes the grand total.
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums(uint64_t* __restrict__ sums, uint64_t n_tiles,
                                                          uint64_t* __restrict__ d_total) {
    uint64_t carry = 0;
    for (uint64_t b0 = 0; b0 < n_tiles; b0 += SCAN_THREADS) {
        uint64_t i = b0 + threadIdx.x;
        uint64_t v = i < n_tiles ? sums[i] : 0;
        uint64_t tot;
        uint64_t inc = block_incl_scan(v, &tot);
        if (i < n_tiles) sums[i] = carry + inc - v;
        carry += tot;
    }
    if (threadIdx.x == 0 && d_total) *d_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const uint64_t* in, uint64_t* out,
                                                           uint64_t n, const uint64_t* __restrict__ sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t x[SCAN_ITEMS];
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        x[i] = base + i < n ? in[base + i] : 0;
        v += x[i];
    }
    uint64_t tot;
    uint64_t inc = block_incl_scan(v, &tot);
    uint64_t run = sums[blockIdx.x] + inc - v;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += x[i];
    }
}

}  // namespace

int hmse_exclusive_scan_u64(hmse_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total,
                            cudaStream_t stream) {
    uint64_t n_tiles = div_up64(n, SCAN_TILE);
    if (n_tiles == 0) n_tiles = 1;
    HMSE_SCRATCH(ctx, sums, uint64_t*, SLOT_SCAN, n_tiles * sizeof(uint64_t));
    KL(ctx);
    scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS, 0, stream>>>(d_in, n, sums);
    KL(ctx);
    scan_sums<<<1, SCAN_THREADS, 0, stream>>>(sums, n_tiles, d_total);
    KL(ctx);
    scan_apply<<<(unsigned)n_tiles, SCAN_THREADS, 0, stream>>>(d_in, d_out, n, sums);
    HMSE_LAUNCH_CHECK(ctx);
    return HMSE_OK;
}
