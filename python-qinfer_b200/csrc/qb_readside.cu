// Read-side estimators evaluated where the cloud lives:
//   qb_weights_entropy    ParticleDistribution.est_entropy (distributions.py:457-465): -sum_{w > 0} w log w
//   qb_weight_mass_hist / qb_weights_select
//                         est_credible_region (distributions.py:558-614) without sorting and without downloading the
//                         cloud.  The region is "the K heaviest particles", K = 1 + #{k : cumsum of the weights in
//                         descending order <= level}: it is determined by a threshold weight tau.  Non-negative
//                         doubles order like their bit patterns, so tau is found by radix SELECTION on the 64-bit
//                         patterns of the normalised weights: six passes, each a histogram (particle count and weight
//                         MASS per 11-bit digit, restricted to the prefix fixed so far); the host walks the 2048
//                         buckets from the top and descends into the one where the accumulated mass crosses `level`.
//                         Then two flag passes (w > tau, w == tau) + the ordered compaction give the member indices;
//                         only the members' rows and weights travel to the host.
// Compiled with --fmad=false.
#include "qb_common.cuh"

namespace qb {

constexpr int RS_THREADS = 256;
constexpr int HIST_BITS = 11;
constexpr int HIST_BUCKETS = 1 << HIST_BITS;

__device__ __forceinline__ unsigned long long weight_bits(double v) {
    // negative, NaN and -0.0 weights carry no mass: pattern 0 (they sort below every positive weight)
    return (v > 0.0) ? static_cast<unsigned long long>(__double_as_longlong(v)) : 0ull;
}

__global__ void __launch_bounds__(RS_THREADS) entropy_kernel(const double* __restrict__ w,
                                                             const double* __restrict__ stats, int64_t n,
                                                             double* __restrict__ partials) {
    __shared__ double red[RS_THREADS / 32];
    const double inv = stats[QB_STAT_INV_NORM];
    double s = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = ldg_stream(w + i) * inv;
        if (v > 0.0) s += log(v) * v;                       // distributions.py:464-465
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < RS_THREADS / 32; ++k) s += red[k];
        partials[blockIdx.x] = s;
    }
}

__global__ void entropy_finish_kernel(const double* __restrict__ partials, int nblocks, double* out) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[b];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = -s;
}

// mass[b] += v, count[b] += 1 for the weights whose bits above (shift + nbits) equal `prefix`; b = digit at `shift`
__global__ void __launch_bounds__(RS_THREADS) mass_hist_kernel(const double* __restrict__ w,
                                                               const double* __restrict__ stats, int64_t n, int shift,
                                                               int nbits, unsigned long long prefix,
                                                               double* __restrict__ mass,
                                                               unsigned long long* __restrict__ count) {
    __shared__ double smass[HIST_BUCKETS];
    __shared__ unsigned int scount[HIST_BUCKETS];
    for (int b = threadIdx.x; b < HIST_BUCKETS; b += RS_THREADS) {
        smass[b] = 0.0;
        scount[b] = 0u;
    }
    __syncthreads();
    const double inv = stats[QB_STAT_INV_NORM];
    const int above = shift + nbits;
    const unsigned long long mask = (1ull << nbits) - 1ull;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = ldg_stream(w + i) * inv;
        const unsigned long long bits = weight_bits(v);
        if (bits == 0ull) continue;
        if (above < 64 && (bits >> above) != prefix) continue;
        const int b = static_cast<int>((bits >> shift) & mask);
        atomicAdd(smass + b, v);
        atomicAdd(scount + b, 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < HIST_BUCKETS; b += RS_THREADS) {
        if (scount[b]) {
            atomicAdd(mass + b, smass[b]);
            atomicAdd(count + b, static_cast<unsigned long long>(scount[b]));
        }
    }
}

// flags[i] = 1 iff bits(w_i) > tau (mode 0) / == tau (mode 1)
__global__ void __launch_bounds__(RS_THREADS) select_kernel(const double* __restrict__ w,
                                                            const double* __restrict__ stats, int64_t n,
                                                            unsigned long long tau, int mode,
                                                            uint8_t* __restrict__ flags) {
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long bits = weight_bits(ldg_stream(w + i) * inv);
        flags[i] = (mode == 0) ? (bits > tau ? 1 : 0) : ((bits == tau && bits != 0ull) ? 1 : 0);
    }
}

static int rs_grid(int64_t n, int per_sm) {
    int64_t want = (n + RS_THREADS - 1) / RS_THREADS;
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (want > cap) want = cap;
    return static_cast<int>(want < 1 ? 1 : want);
}

}  // namespace qb

using namespace qb;

extern "C" size_t qb_readside_workspace_bytes(void) {
    return 256 + static_cast<size_t>(HIST_BUCKETS) * 16 + 8 * 4096;
}

extern "C" int qb_weights_entropy(const double* d_w, const double* d_stats, int64_t n, double* d_out, void* d_ws,
                                  size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_w && d_stats && d_out && d_ws && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_weights_entropy: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_readside_workspace_bytes(), QB_ERR_WORKSPACE, "qb_weights_entropy: workspace too small");
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    int grid = rs_grid(n, 8);
    if (grid > 4096) grid = 4096;
    cudaStream_t st = as_stream(stream);
    entropy_kernel<<<grid, RS_THREADS, 0, st>>>(d_w, d_stats, n, partials);
    QB_CUDA_CHECK(cudaGetLastError());
    entropy_finish_kernel<<<1, 32, 0, st>>>(partials, grid, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_weight_mass_hist(const double* d_w, const double* d_stats, int64_t n, int32_t shift, int32_t nbits,
                                   uint64_t prefix, double* d_mass, uint64_t* d_count, void* stream) {
    QB_REQUIRE(d_w && d_stats && d_mass && d_count && n >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_weight_mass_hist: bad arguments");
    QB_REQUIRE(nbits >= 1 && nbits <= HIST_BITS && shift >= 0 && shift + nbits <= 64, QB_ERR_INVALID_ARGUMENT,
               "qb_weight_mass_hist: digit [%d, %d) outside the 64-bit pattern / wider than %d bits", shift,
               shift + nbits, HIST_BITS);
    cudaStream_t st = as_stream(stream);
    QB_CUDA_CHECK(cudaMemsetAsync(d_mass, 0, sizeof(double) * HIST_BUCKETS, st));
    QB_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t) * HIST_BUCKETS, st));
    mass_hist_kernel<<<rs_grid(n, 4), RS_THREADS, 0, st>>>(d_w, d_stats, n, shift, nbits, prefix, d_mass,
                                                           reinterpret_cast<unsigned long long*>(d_count));
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_weights_select(const double* d_w, const double* d_stats, int64_t n, uint64_t tau_bits, int32_t mode,
                                 uint8_t* d_flags, void* stream) {
    QB_REQUIRE(d_w && d_stats && d_flags && n >= 1 && (mode == 0 || mode == 1), QB_ERR_INVALID_ARGUMENT,
               "qb_weights_select: bad arguments");
    select_kernel<<<rs_grid(n, 8), RS_THREADS, 0, as_stream(stream)>>>(d_w, d_stats, n, tau_bits, mode, d_flags);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
