// Counter-based device RNG for the throughput mode of the resampler
// (Philox4x32-10, Salmon et al. SC'11).  Parity mode does not use this: it draws
// from NumPy's legacy MT19937 stream on the host exactly as the reference does
// (resamplers.py:319,332) and uploads the variates.
//
// Element i of stream (seed, offset) uses counter (offset + i/2), key = seed:
//   uniform: ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two of the four 32-bit
//            words (same 53-bit construction as NumPy's random_sample)
//   normal : Box-Muller on the two uniforms of one counter value.
#include "qb_philox.cuh"

namespace qb {

template <bool NORMAL>
__global__ void __launch_bounds__(256) rng_kernel(double* __restrict__ out, int64_t n, uint64_t seed,
                                                  uint64_t offset) {
    const int64_t npairs = (n + 1) / 2;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < npairs; p += stride) {
        const uint64_t ctr = offset + static_cast<uint64_t>(p);
        double a, b;
        if (NORMAL)
            philox_normal_pair(seed, ctr, a, b);
        else
            philox_uniform_pair(seed, ctr, a, b);
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
