// Per-SM issue rates of the integer / warp-collective instructions parse_kernel leans on (B200, sm_100a).
// One CTA of 1024 threads per SM (8 warps per scheduler), independent chains, clock64 around the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench/ubench.cu && ./ubench
// Output: warp instructions per cycle per SM for each op (4.0 = one per scheduler per cycle).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2000, CH = 8;

#define KERNEL(NAME, INIT, BODY)                                                              \
    __global__ void __launch_bounds__(1024) k_##NAME(uint32_t* out, long long* cyc, uint32_t seed) { \
        __shared__ uint32_t sh[4096];                                                         \
        const unsigned lane = threadIdx.x & 31;                                               \
        uint32_t x[CH];                                                                       \
        for (int c = 0; c < CH; c++) x[c] = seed * (threadIdx.x + 1) + c * 0x9E3779B1u;     \
        for (int i = threadIdx.x; i < 4096; i += 1024) sh[i] = i * seed;                     \
        INIT;                                                                                 \
        __syncthreads();                                                                      \
        const long long t0 = clock64();                                                       \
        for (int it = 0; it < ITERS; it++) {                                                  \
            _Pragma("unroll") for (int c = 0; c < CH; c++) { BODY; }                          \
        }                                                                                     \
        const long long t1 = clock64();                                                       \
        uint32_t acc = 0;                                                                     \
        for (int c = 0; c < CH; c++) acc ^= x[c];                                             \
        out[blockIdx.x * 1024 + threadIdx.x] = acc + sh[threadIdx.x] + lane;                  \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                      \
    }

KERNEL(lop3, , x[c] = (x[c] ^ seed) & (x[c] >> 3 | seed))   // 2: SHF + LOP3
KERNEL(iadd, , x[c] = x[c] + seed + c)
KERNEL(imad, , x[c] = x[c] * seed + c)
KERNEL(popc, , x[c] = __popc(x[c]) + seed)                   // POPC + IADD
KERNEL(ffs, , x[c] = __ffs(x[c]) + seed)
KERNEL(brev, , x[c] = __brev(x[c]) + seed)
KERNEL(prmt, , x[c] = __byte_perm(x[c], seed, 0x5410 + c))
KERNEL(shf, , x[c] = __funnelshift_r(x[c], seed, c + 1))
KERNEL(vote, , x[c] = __ballot_sync(0xffffffffu, x[c] > seed) + c)    // ISETP + VOTE + IADD
KERNEL(shfl, , x[c] = __shfl_up_sync(0xffffffffu, x[c], 1) + c)
KERNEL(match, , x[c] = __match_any_sync(0xffffffffu, x[c] & 7) + c)
KERNEL(redux, , x[c] = __reduce_add_sync(0xffffffffu, x[c]) + c)
KERNEL(lds, , x[c] = sh[(x[c] + lane) & 4095 & ~31u | lane] + c)     // conflict-free LDS + address ops
KERNEL(sts, , sh[(c * 32 + lane + threadIdx.x) & 4095] = x[c]; x[c] += seed)
KERNEL(atoms, , atomicMax(&sh[(c * 32 + lane + (threadIdx.x & ~31u)) & 4095], x[c]); x[c] += seed)
KERNEL(vimnmx, , x[c] = max(x[c] ^ seed, (uint32_t)c))
KERNEL(sel, , x[c] = x[c] > seed ? x[c] - 1 : seed)
KERNEL(mulhi, , x[c] = __umulhi(x[c], seed) + c)
KERNEL(rotmul, , { const uint32_t lo = x[c] * 0x02000000u; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(x[c]) : "r"(x[c]), "r"(0x02000000u), "r"(lo)); })
KERNEL(rotshf, , x[c] = __funnelshift_r(x[c], x[c], 7) ^ seed)

struct T { const char* name; void (*fn)(uint32_t*, long long*, uint32_t); int per_body; };

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * 1024 * 4);
    cudaMallocManaged(&cyc, sms * 8);
    T tests[] = {{"shf+lop3", k_lop3, 2}, {"iadd3", k_iadd, 1}, {"imad", k_imad, 1}, {"popc+iadd", k_popc, 2}, {"ffs(brev+flo)+iadd", k_ffs, 3},
                 {"brev+iadd", k_brev, 2}, {"prmt", k_prmt, 1}, {"shf", k_shf, 1}, {"isetp+vote+iadd", k_vote, 3}, {"shfl+iadd", k_shfl, 2},
                 {"match.any+lop+iadd", k_match, 3}, {"redux+iadd", k_redux, 2}, {"lds(+3 addr)", k_lds, 4}, {"sts+iadd", k_sts, 2},
                 {"atoms.max+iadd", k_atoms, 2}, {"lop3+vimnmx", k_vimnmx, 2}, {"isetp+sel(+iadd)", k_sel, 3}, {"imad.hi+iadd", k_mulhi, 2}, {"rot = imad + imad.hi", k_rotmul, 2}, {"rot = shf (+lop3)", k_rotshf, 2}};
    for (auto& t : tests) {
        t.fn<<<sms, 1024>>>(out, cyc, 12345u);
        cudaDeviceSynchronize();
        t.fn<<<sms, 1024>>>(out, cyc, 12345u);
        cudaError_t e = cudaDeviceSynchronize();
        long long mx = 0;
        for (int i = 0; i < sms; i++) mx = cyc[i] > mx ? cyc[i] : mx;
        const double bodies = 32.0 * ITERS * CH;   // warp-level bodies per SM
        printf("%-22s %s cycles %lld  cycles per body per scheduler %.2f  (source-level ops per body ~%d)\n", t.name, cudaGetErrorString(e), mx,
               mx / (bodies / 4.0), t.per_body);
    }
    return 0;
}
