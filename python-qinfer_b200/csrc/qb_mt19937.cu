// NumPy's legacy MT19937 stream on the device — parity mode at scale.
//
// The reference draws its resampling variates from the GLOBAL legacy np.random stream: np.random.random((n,))
// (resamplers.py:319) and np.random.randn(n_rvs, k) per retry (resamplers.py:332).  Reproducing the resample
// indices bit for bit therefore needs the very same uniforms; generating 10^7 of them on the host costs ~70 ms
// (and ~300 ms for the normals), two orders of magnitude more than the rest of a resample.  These kernels continue
// the stream from the state np.random.get_state() returns and hand back the state to np.random.set_state():
//
//   raw stream   R[0..623] = key,  R[p] = R[p-227] ^ (y >> 1) ^ (y & 1 ? 0x9908b0df : 0),
//                y = (R[p-624] & 0x80000000) | (R[p-623] & 0x7fffffff)        (the MT19937 recurrence, unrolled
//                in time: position p only needs p-227, p-623, p-624, so 227 consecutive positions are independent)
//                -> one CTA advances 227 words per step out of a shared-memory ring and streams them to HBM;
//   uniforms     mt19937_next_double: ((T(R[q]) >> 5) * 2^26 + (T(R[q+1]) >> 6)) / 2^53, T = tempering   (bit-exact)
//   normals      legacy_gauss (polar method with a cached second variate): candidates are consecutive 4-word groups;
//                acceptance (r2 < 1 && r2 != 0) uses only exactly rounded operations, so the accepted set, the
//                number of words consumed and the final generator state are exact; the variates themselves go
//                through log(), which is within 1 ulp of glibc's (measured: identical for > 99.9 % of the draws).
#include <cstring>
#include "qb_common.cuh"

namespace qb {

constexpr int MT_N = 624, MT_M = 397, MT_LAG = MT_N - MT_M;  // 227
constexpr int MT_RING = 2048;

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// One CTA, 227 active lanes.  Lane t produces positions p = p0 + t and p + 227 per barrier: the second only needs
// R[p] (its own, still in a register) and R[p - 397], R[p - 396], which are older than p0 for every t < 227.
__global__ void __launch_bounds__(256, 1) mt_raw_kernel(uint32_t* __restrict__ raw, int64_t nwords_total) {
    __shared__ uint32_t ring[MT_RING];
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += 256) ring[i] = raw[i];
    __syncthreads();
    for (int64_t p0 = MT_N; p0 < nwords_total; p0 += 2 * MT_LAG) {
        if (tid < MT_LAG) {
            const int64_t p = p0 + tid;
            const uint32_t v0 = mt_twist(ring[(p - MT_N) & (MT_RING - 1)], ring[(p - MT_N + 1) & (MT_RING - 1)],
                                         ring[(p - MT_LAG) & (MT_RING - 1)]);
            const uint32_t v1 = mt_twist(ring[(p - MT_M) & (MT_RING - 1)], ring[(p - MT_M + 1) & (MT_RING - 1)], v0);
            ring[p & (MT_RING - 1)] = v0;
            ring[(p + MT_LAG) & (MT_RING - 1)] = v1;
            if (p < nwords_total) raw[p] = v0;
            if (p + MT_LAG < nwords_total) raw[p + MT_LAG] = v1;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ double mt_double(const uint32_t* __restrict__ raw, int64_t q) {
    const uint32_t a = mt_temper(raw[q]) >> 5, b = mt_temper(raw[q + 1]) >> 6;
    return (static_cast<double>(a) * 67108864.0 + static_cast<double>(b)) / 9007199254740992.0;
}

__global__ void __launch_bounds__(256) mt_uniform_kernel(const uint32_t* __restrict__ raw, int64_t pos, int64_t n,
                                                         double* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = mt_double(raw, pos + 2 * i);
}

// candidate c uses words [pos + 4c, pos + 4c + 4)
__device__ __forceinline__ bool gauss_candidate(const uint32_t* __restrict__ raw, int64_t pos, int64_t c, double& x1,
                                                double& x2, double& r2) {
    x1 = 2.0 * mt_double(raw, pos + 4 * c) - 1.0;
    x2 = 2.0 * mt_double(raw, pos + 4 * c + 2) - 1.0;
    r2 = x1 * x1 + x2 * x2;  // --fmad=false: two exactly rounded products and a sum, as the C reference
    return !(r2 >= 1.0 || r2 == 0.0);
}

__global__ void __launch_bounds__(256) mt_gauss_flag_kernel(const uint32_t* __restrict__ raw, int64_t pos,
                                                            int64_t ncand, uint8_t* __restrict__ flags) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < ncand; c += stride) {
        double x1, x2, r2;
        flags[c] = gauss_candidate(raw, pos, c, x1, x2, r2) ? 1 : 0;
    }
}

// accepted pair a -> out[off + 2a] = f * x2, out[off + 2a + 1] = f * x1 (legacy_gauss returns f*x2 first, caches f*x1)
__global__ void __launch_bounds__(256) mt_gauss_write_kernel(const uint32_t* __restrict__ raw, int64_t pos,
                                                             const int64_t* __restrict__ idxs, int64_t npairs,
                                                             double* __restrict__ out, int64_t off, int64_t m,
                                                             double* __restrict__ last_cached) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t a = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; a < npairs; a += stride) {
        double x1, x2, r2;
        gauss_candidate(raw, pos, idxs[a], x1, x2, r2);
        const double f = sqrt(-2.0 * log(r2) / r2);
        const int64_t o = off + 2 * a;
        out[o] = f * x2;
        if (o + 1 < m)
            out[o + 1] = f * x1;
        else
            *last_cached = f * x1;  // odd request: the second variate stays cached in the generator state
    }
}

static int grid_cap(int64_t want, int per_sm) {
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

// state after consuming words up to (not including) stream index q
static void final_block(int64_t q, int64_t* block, int32_t* pos_out) {
    if (q % MT_N == 0 && q > 0) {
        *block = q / MT_N - 1;
        *pos_out = MT_N;  // NumPy regenerates lazily: a full block consumed leaves pos == 624
    } else {
        *block = q / MT_N;
        *pos_out = static_cast<int32_t>(q - MT_N * (*block));
    }
}

}  // namespace qb

using namespace qb;

extern "C" int qb_compact_invalid(const uint8_t*, int64_t, int64_t*, int64_t*, void*, size_t, void*);
extern "C" size_t qb_compact_workspace_bytes(int64_t);

static int64_t gauss_candidates_for(int64_t npairs) {
    return static_cast<int64_t>(static_cast<double>(npairs) / 0.7853981633974483 * 1.01) + 4096;
}

extern "C" size_t qb_mt19937_workspace_bytes(int64_t n_uniform, int64_t n_normal) {
    int64_t words = 2 * n_uniform + 2 * MT_N + MT_LAG + 1024;
    size_t extra = 0;
    if (n_normal > 0) {
        const int64_t ncand = 2 * gauss_candidates_for((n_normal + 1) / 2);  // room for one retry at twice the size
        const int64_t w2 = 4 * ncand + 2 * MT_N + MT_LAG + 1024;
        if (w2 > words) words = w2;
        extra = static_cast<size_t>(ncand) * (1 + 8) + qb_compact_workspace_bytes(ncand) + 1024;
    }
    return static_cast<size_t>(words) * sizeof(uint32_t) + extra + 1024;
}

static int mt_generate(const uint32_t* h_key, uint32_t* d_raw, int64_t q_max, cudaStream_t st) {
    QB_CUDA_CHECK(cudaMemcpyAsync(d_raw, h_key, MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const int64_t total = q_max + MT_N + 1;  // enough to cut the final 624-word state block out of the stream
    mt_raw_kernel<<<1, 256, 0, st>>>(d_raw, total);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_mt19937_uniform(const uint32_t* h_key, int32_t pos, int64_t n, double* d_out, uint32_t* h_key_out,
                                  int32_t* pos_out, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(h_key && d_out && h_key_out && pos_out && d_ws && n >= 1 && pos >= 0 && pos <= MT_N,
               QB_ERR_INVALID_ARGUMENT, "qb_mt19937_uniform: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_mt19937_workspace_bytes(n, 0), QB_ERR_WORKSPACE, "qb_mt19937_uniform: workspace too small");
    cudaStream_t st = as_stream(stream);
    uint32_t* raw = reinterpret_cast<uint32_t*>(d_ws);
    const int64_t q = static_cast<int64_t>(pos) + 2 * n;
    int rc = mt_generate(h_key, raw, q, st);
    if (rc != QB_OK) return rc;
    mt_uniform_kernel<<<grid_cap((n + 255) / 256, 8), 256, 0, st>>>(raw, pos, n, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    int64_t block;
    final_block(q, &block, pos_out);
    QB_CUDA_CHECK(cudaMemcpyAsync(h_key_out, raw + block * MT_N, MT_N * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

extern "C" int qb_mt19937_normal(const uint32_t* h_key, int32_t pos, int32_t has_gauss, double cached, int64_t m,
                                 double* d_out, uint32_t* h_key_out, int32_t* pos_out, int32_t* has_gauss_out,
                                 double* cached_out, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(h_key && d_out && h_key_out && pos_out && has_gauss_out && cached_out && d_ws && m >= 1 && pos >= 0 &&
                   pos <= MT_N,
               QB_ERR_INVALID_ARGUMENT, "qb_mt19937_normal: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_mt19937_workspace_bytes(0, m), QB_ERR_WORKSPACE, "qb_mt19937_normal: workspace too small");
    cudaStream_t st = as_stream(stream);
    int64_t off = 0;
    if (has_gauss) {  // the cached second variate of an earlier call comes out first
        QB_CUDA_CHECK(cudaMemcpyAsync(d_out, &cached, sizeof(double), cudaMemcpyHostToDevice, st));
        off = 1;
    }
    const int64_t need = m - off;
    const int64_t npairs = (need + 1) / 2;
    if (npairs == 0) {
        memcpy(h_key_out, h_key, MT_N * sizeof(uint32_t));
        *pos_out = pos;
        *has_gauss_out = 0;
        *cached_out = 0.0;
        QB_CUDA_CHECK(cudaStreamSynchronize(st));
        return QB_OK;
    }
    const int64_t ncand_max = 2 * gauss_candidates_for(npairs);
    unsigned char* base = reinterpret_cast<unsigned char*>(d_ws);
    uint32_t* raw = reinterpret_cast<uint32_t*>(base);
    const size_t raw_bytes = ((static_cast<size_t>(4 * ncand_max + 2 * MT_N + MT_LAG + 1024) * 4 + 255) / 256) * 256;
    int64_t* idxs = reinterpret_cast<int64_t*>(base + raw_bytes);
    uint8_t* flags = reinterpret_cast<uint8_t*>(base + raw_bytes + static_cast<size_t>(ncand_max) * 8);
    unsigned char* cws = base + raw_bytes + static_cast<size_t>(ncand_max) * 9;
    cws += (256 - (reinterpret_cast<uintptr_t>(cws) & 255)) & 255;
    int64_t* d_count = reinterpret_cast<int64_t*>(cws);
    double* d_cached = reinterpret_cast<double*>(cws + 64);
    void* compact_ws = cws + 256;
    const size_t compact_bytes = qb_compact_workspace_bytes(ncand_max);

    int64_t ncand = gauss_candidates_for(npairs);
    int64_t h_count = 0, last_idx = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = mt_generate(h_key, raw, static_cast<int64_t>(pos) + 4 * ncand, st);
        if (rc != QB_OK) return rc;
        mt_gauss_flag_kernel<<<grid_cap((ncand + 255) / 256, 8), 256, 0, st>>>(raw, pos, ncand, flags);
        QB_CUDA_CHECK(cudaGetLastError());
        rc = qb_compact_invalid(flags, ncand, idxs, d_count, compact_ws, compact_bytes, stream);
        if (rc != QB_OK) return rc;
        QB_CUDA_CHECK(cudaMemcpyAsync(&h_count, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        QB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_count >= npairs) break;
        ncand = ncand_max;  // one retry with twice the candidates (probability of needing it ~ e^-100)
    }
    QB_REQUIRE(h_count >= npairs, QB_ERR_CUDA, "qb_mt19937_normal: rejection sampler ran out of candidates");
    QB_CUDA_CHECK(cudaMemcpyAsync(&last_idx, idxs + (npairs - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    mt_gauss_write_kernel<<<grid_cap((npairs + 255) / 256, 8), 256, 0, st>>>(raw, pos, idxs, npairs, d_out, off, m,
                                                                              d_cached);
    QB_CUDA_CHECK(cudaGetLastError());
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    const int64_t q = static_cast<int64_t>(pos) + 4 * (last_idx + 1);
    int64_t block;
    final_block(q, &block, pos_out);
    QB_CUDA_CHECK(cudaMemcpyAsync(h_key_out, raw + block * MT_N, MT_N * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    *has_gauss_out = (need % 2 == 1) ? 1 : 0;
    *cached_out = 0.0;
    if (*has_gauss_out) QB_CUDA_CHECK(cudaMemcpyAsync(cached_out, d_cached, sizeof(double), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}
