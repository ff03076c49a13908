// Sharded particle cloud: one contiguous slab of particles per GPU.
//   * peer mailboxes + CUDA IPC plumbing for the in-kernel all-reduce of the fused update
//     (qb_update.cu: peer_allreduce3);
//   * the routing kernels of the resample exchange: classify each uniform draw to the shard that
//     owns that part of the global CDF, bucket the owner-local coordinates for the all-to-all, and
//     gather the rows an owner sends back.
#include <cstring>
#include "qb_common.cuh"

namespace qb {

struct BoundsArg {
    double b[QB_MAX_RANKS + 1];
    long long start[QB_MAX_RANKS];
};

__device__ __forceinline__ int owner_of(const BoundsArg& a, int G, double u) {
    // last shard r with bounds[r] <= u  (bounds ascending, bounds[0] = 0); draws beyond the total go to the last
    int r = 0;
#pragma unroll 4
    for (int k = 1; k < G; ++k) r += (a.b[k] <= u) ? 1 : 0;
    return r;
}

__global__ void __launch_bounds__(256) classify_kernel(const double* __restrict__ u, int64_t n,
                                                       const __grid_constant__ BoundsArg a, int G,
                                                       int32_t* __restrict__ owner,
                                                       unsigned long long* __restrict__ counts) {
    __shared__ unsigned int hist[QB_MAX_RANKS];
    if (threadIdx.x < QB_MAX_RANKS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = owner_of(a, G, u[i]);
        owner[i] = r;
        atomicAdd(&hist[r], 1u);
    }
    __syncthreads();
    if (threadIdx.x < G && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(hist[threadIdx.x]));
}

// Two-level slot claiming: a block counts its draws per owner in shared memory, claims one contiguous range per
// owner with ONE global atomic each, and hands out positions inside the range by warp ballot + per-warp offsets.
__global__ void __launch_bounds__(256) bucket_kernel(const double* __restrict__ u, const int32_t* __restrict__ owner,
                                                     int64_t n, const __grid_constant__ BoundsArg a, int G,
                                                     unsigned long long* __restrict__ cursor,
                                                     double* __restrict__ req, int64_t* __restrict__ perm) {
    __shared__ unsigned int warp_cnt[8][QB_MAX_RANKS];   // per-warp count per owner, then exclusive offsets
    __shared__ unsigned long long block_base[QB_MAX_RANKS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t per_block = static_cast<int64_t>(blockDim.x);
    const int64_t ntiles = (n + per_block - 1) / per_block;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t i = t * per_block + threadIdx.x;
        const bool live = i < n;
        const int r = live ? owner[i] : -1;
        unsigned int my_rank_in_warp = 0;
        for (int g = 0; g < G; ++g) {
            const unsigned int m = __ballot_sync(0xffffffffu, r == g);
            if (lane == 0) warp_cnt[wid][g] = __popc(m);
            if (r == g) my_rank_in_warp = __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < G) {  // exclusive scan over the 8 warps + one global atomic per owner
            const int g = threadIdx.x;
            unsigned int run = 0;
            for (int w = 0; w < 8; ++w) {
                const unsigned int c = warp_cnt[w][g];
                warp_cnt[w][g] = run;
                run += c;
            }
            block_base[g] = run ? atomicAdd(&cursor[g], static_cast<unsigned long long>(run)) : 0ULL;
        }
        __syncthreads();
        if (live) {
            const long long pos = a.start[r] + static_cast<long long>(block_base[r]) + warp_cnt[wid][r] + my_rank_in_warp;
            req[pos] = u[i] - a.b[r];
            perm[i] = pos;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ x, int d,
                                                          const int64_t* __restrict__ js, int64_t n,
                                                          double* __restrict__ out) {
    const int64_t total = n * d;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / d;
        const int c = static_cast<int>(e - i * d);
        out[e] = __ldg(x + js[i] * d + c);
    }
}

static int grid_cap(int64_t want, int per_sm) {
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

}  // namespace qb

using namespace qb;

extern "C" int qb_mailbox_create(int32_t n_ranks, double** d_mailbox) {
    QB_REQUIRE(d_mailbox && n_ranks >= 1 && n_ranks <= QB_MAX_RANKS, QB_ERR_INVALID_ARGUMENT,
               "qb_mailbox_create: bad arguments");
    const size_t bytes = static_cast<size_t>(2) * n_ranks * QB_MAILBOX_ROW * sizeof(double);
    void* p = nullptr;
    QB_CUDA_CHECK(cudaMalloc(&p, bytes));
    QB_CUDA_CHECK(cudaMemset(p, 0, bytes));
    QB_CUDA_CHECK(cudaDeviceSynchronize());
    *d_mailbox = static_cast<double*>(p);
    return QB_OK;
}

extern "C" int qb_mailbox_destroy(double* d_mailbox) {
    if (d_mailbox) QB_CUDA_CHECK(cudaFree(d_mailbox));
    return QB_OK;
}

extern "C" int qb_ipc_get_handle(const void* d_ptr, unsigned char handle[QB_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == QB_IPC_HANDLE_BYTES, "IPC handle size");
    QB_REQUIRE(d_ptr && handle, QB_ERR_INVALID_ARGUMENT, "qb_ipc_get_handle: bad arguments");
    cudaIpcMemHandle_t h;
    QB_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, sizeof(h));
    return QB_OK;
}

extern "C" int qb_ipc_open_handle(const unsigned char handle[QB_IPC_HANDLE_BYTES], void** d_ptr) {
    QB_REQUIRE(d_ptr && handle, QB_ERR_INVALID_ARGUMENT, "qb_ipc_open_handle: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    QB_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QB_OK;
}

extern "C" int qb_ipc_close_handle(void* d_ptr) {
    if (d_ptr) QB_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return QB_OK;
}

static int fill_bounds(BoundsArg& a, const double* h_bounds, const int64_t* h_start, int n_ranks) {
    for (int r = 0; r <= QB_MAX_RANKS; ++r) a.b[r] = (r <= n_ranks) ? h_bounds[r] : INFINITY;
    for (int r = 0; r < QB_MAX_RANKS; ++r) a.start[r] = (h_start && r < n_ranks) ? h_start[r] : 0;
    return QB_OK;
}

extern "C" int qb_shard_classify(const double* d_u, int64_t n, const double* h_bounds, int32_t n_ranks,
                                 int32_t* d_owner, int64_t* d_counts, void* stream) {
    QB_REQUIRE(d_u && h_bounds && d_owner && d_counts && n >= 1 && n_ranks >= 1 && n_ranks <= QB_MAX_RANKS,
               QB_ERR_INVALID_ARGUMENT, "qb_shard_classify: bad arguments");
    BoundsArg a;
    fill_bounds(a, h_bounds, nullptr, n_ranks);
    cudaStream_t st = as_stream(stream);
    QB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, sizeof(int64_t) * n_ranks, st));
    classify_kernel<<<grid_cap((n + 255) / 256, 8), 256, 0, st>>>(d_u, n, a, n_ranks, d_owner,
                                                                  reinterpret_cast<unsigned long long*>(d_counts));
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_shard_bucket(const double* d_u, const int32_t* d_owner, int64_t n, const double* h_bounds,
                               const int64_t* h_bucket_start, int32_t n_ranks, int64_t* d_cursor, double* d_req,
                               int64_t* d_perm, void* stream) {
    QB_REQUIRE(d_u && d_owner && h_bounds && h_bucket_start && d_cursor && d_req && d_perm && n >= 1 &&
                   n_ranks >= 1 && n_ranks <= QB_MAX_RANKS,
               QB_ERR_INVALID_ARGUMENT, "qb_shard_bucket: bad arguments");
    BoundsArg a;
    fill_bounds(a, h_bounds, h_bucket_start, n_ranks);
    cudaStream_t st = as_stream(stream);
    QB_CUDA_CHECK(cudaMemsetAsync(d_cursor, 0, sizeof(int64_t) * n_ranks, st));
    bucket_kernel<<<grid_cap((n + 255) / 256, 8), 256, 0, st>>>(d_u, d_owner, n, a, n_ranks,
                                                                reinterpret_cast<unsigned long long*>(d_cursor),
                                                                d_req, d_perm);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_gather_rows(const double* d_x, int32_t d, const int64_t* d_js, int64_t n, double* d_out,
                              void* stream) {
    QB_REQUIRE(d_x && d_js && d_out && d >= 1, QB_ERR_INVALID_ARGUMENT, "qb_gather_rows: bad arguments");
    if (n == 0) return QB_OK;
    gather_rows_kernel<<<grid_cap((n * d + 255) / 256, 8), 256, 0, as_stream(stream)>>>(d_x, d, d_js, n, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
