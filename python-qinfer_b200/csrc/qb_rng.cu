// Counter-based device RNG for the throughput mode of the resampler
// (Philox4x32-10, Salmon et al. SC'11).  Parity mode does not use this: it draws
// from NumPy's legacy MT19937 stream on the host exactly as the reference does
// (resamplers.py:319,332) and uploads the variates.
//
// Element i of stream (seed, offset) uses counter (offset + i/2), key = seed:
//   uniform: ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two of the four 32-bit
//            words (same 53-bit construction as NumPy's random_sample)
//   normal : Box-Muller on the two uniforms of one counter value.
#include "qb_common.cuh"

namespace qb {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return (static_cast<double>(a >> 5) * 67108864.0 + static_cast<double>(b >> 6)) * (1.0 / 9007199254740992.0);
}

template <bool NORMAL>
__global__ void __launch_bounds__(256) rng_kernel(double* __restrict__ out, int64_t n, uint64_t seed,
                                                  uint64_t offset) {
    const int64_t npairs = (n + 1) / 2;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < npairs; p += stride) {
        const uint64_t ctr = offset + static_cast<uint64_t>(p);
        uint32_t r[4];
        philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u,
                      static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
        double a = u53(r[0], r[1]);
        double b = u53(r[2], r[3]);
        if (NORMAL) {
            const double rad = sqrt(-2.0 * log(1.0 - a));  // 1 - a in (0, 1]
            double sn, cs;
            sincospi(2.0 * b, &sn, &cs);
            a = rad * cs;
            b = rad * sn;
        }
        out[2 * p] = a;
        if (2 * p + 1 < n) out[2 * p + 1] = b;
    }
}

}  // namespace qb

using namespace qb;

static int rng_launch(double* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream, bool normal) {
    QB_REQUIRE(d_out && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_rng: bad arguments");
    int64_t want = ((n + 1) / 2 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (want > cap) want = cap;
    const int grid = static_cast<int>(want < 1 ? 1 : want);
    if (normal)
        rng_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(d_out, n, seed, offset);
    else
        rng_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(d_out, n, seed, offset);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_rng_uniform(double* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    return rng_launch(d_out, n, seed, offset, stream, false);
}

extern "C" int qb_rng_normal(double* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    return rng_launch(d_out, n, seed, offset, stream, true);
}
