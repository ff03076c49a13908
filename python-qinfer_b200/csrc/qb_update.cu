// Fused Bayes-update kernel: likelihood x weight multiply x normalisation sum x
// n_ess reduction in ONE launch (smc.py:324-386,413-453).
//
// HBM traffic per particle-update: read x (8d) + read w (8) + write w' (8)
// = 8(d+2) bytes; the division by the normalisation (smc.py:373) is deferred
// as the scalar stats[INV_NORM] that the NEXT pass folds into its weight load.
//
// Data movement: full tiles of the row-major (n, d) particle slab and of the
// weight vector are staged global->shared with 1-D bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx) through a STAGES-deep ring, so every
// SM keeps STAGES x ~16 KB of loads in flight without holding them in
// registers; the fp64 likelihood math then runs out of shared memory and the
// new weights are written back with fully coalesced 8-B streaming stores.
// The ragged last tile (byte count not a multiple of 16) is loaded directly.
#include "qb_models.cuh"

namespace qb {

constexpr int UPD_THREADS = 256;
constexpr int UPD_STAGES = 3;

struct UpdateParams {
    const double* x;
    const double* w_in;
    double* w_out;
    const double* stats_in;
    double* stats_out;
    double* partials;        // [grid][4]
    unsigned int* ticket;    // last-block-done counter (self-resetting)
    int64_t n;
    int32_t tile;            // particles per tile
    int32_t d;
    ModelView mv;
    ExpView ev;
    double meas[QB_MAX_D];
};

__device__ __forceinline__ void block_reduce4(double& s, double& q, double& mn, double& bad, double* red) {
    s = warp_sum(s);
    q = warp_sum(q);
    mn = warp_min(mn);
    bad = warp_sum(bad);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[wid * 4 + 0] = s;
        red[wid * 4 + 1] = q;
        red[wid * 4 + 2] = mn;
        red[wid * 4 + 3] = bad;
    }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        s = (lane < nw) ? red[lane * 4 + 0] : 0.0;
        q = (lane < nw) ? red[lane * 4 + 1] : 0.0;
        mn = (lane < nw) ? red[lane * 4 + 2] : INFINITY;
        bad = (lane < nw) ? red[lane * 4 + 3] : 0.0;
        s = warp_sum(s);
        q = warp_sum(q);
        mn = warp_min(mn);
        bad = warp_sum(bad);
    }
}

// Final deterministic reduction of per-block partials by the last block to finish.
__device__ void finish_stats(const double* partials, int nblocks, double* stats_out, double* red) {
    double s = 0.0, q = 0.0, mn = INFINITY, bad = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        s += partials[b * 4 + 0];
        q += partials[b * 4 + 1];
        mn = fmin(mn, partials[b * 4 + 2]);
        bad += partials[b * 4 + 3];
    }
    __syncthreads();
    block_reduce4(s, q, mn, bad, red);
    if (threadIdx.x == 0) {
        const double eps = 2.220446049250313e-16;  // np.spacing(1), smc.py:370
        stats_out[QB_STAT_NORM] = s;
        stats_out[QB_STAT_SUMSQ] = q;
        stats_out[QB_STAT_MIN] = mn;
        stats_out[QB_STAT_NBAD] = bad;
        stats_out[QB_STAT_INV_NORM] = (fabs(s) < eps) ? 1.0 : 1.0 / s;
        stats_out[QB_STAT_NESS] = (s * s) / q;
    }
}

template <int KIND, bool BINOM>
__global__ void __launch_bounds__(UPD_THREADS) fused_update_kernel(const __grid_constant__ UpdateParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int d = p.d;
    const int tile = p.tile;
    const uint32_t x_bytes = static_cast<uint32_t>(tile) * d * 8u;
    const uint32_t w_bytes = static_cast<uint32_t>(tile) * 8u;
    const uint32_t stage_bytes = x_bytes + w_bytes;  // multiples of 128 by construction (tile % 16 == 0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // UPD_STAGES barriers in the first 128 B
    double* meas_s = reinterpret_cast<double*>(smem_raw + 128);
    unsigned char* ring = smem_raw + 128 + QB_MAX_D * 8;
    __shared__ double red[(UPD_THREADS / 32) * 4];
    __shared__ unsigned int is_last;

    const int tid = threadIdx.x;
    const int64_t ntiles = (p.n + tile - 1) / tile;
    const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (KIND == QB_MODEL_TOMOGRAPHY) {
        for (int c = tid; c < d; c += UPD_THREADS) meas_s[c] = p.meas[c];
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < UPD_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int64_t i) {  // thread 0: start the bulk loads of my i-th tile if it is a full one
        const int64_t t = blockIdx.x + i * gridDim.x;
        const int64_t first = t * tile;
        if (first + tile <= p.n) {
            const int s = static_cast<int>(i % UPD_STAGES);
            unsigned char* dst = ring + static_cast<size_t>(s) * stage_bytes;
            mbar_expect_tx(&bars[s], stage_bytes);
            tma_load_1d(dst, p.x + first * d, x_bytes, &bars[s]);
            tma_load_1d(dst + x_bytes, p.w_in + first, w_bytes, &bars[s]);
        }
    };
    if (tid == 0) {
        for (int64_t i = 0; i < my_tiles && i < UPD_STAGES; ++i) issue(i);
    }

    const double inv_norm = p.stats_in[QB_STAT_INV_NORM];
    double acc_s = 0.0, acc_q = 0.0, acc_min = INFINITY, acc_bad = 0.0;
    const int lane = tid & 31;

    for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = static_cast<int>(i % UPD_STAGES);
        const uint32_t parity = static_cast<uint32_t>((i / UPD_STAGES) & 1);
        const int64_t t = blockIdx.x + i * gridDim.x;
        const int64_t first = t * tile;
        const int cnt = static_cast<int>((p.n - first < tile) ? (p.n - first) : tile);
        double* xs = reinterpret_cast<double*>(ring + static_cast<size_t>(s) * stage_bytes);
        double* ws = reinterpret_cast<double*>(ring + static_cast<size_t>(s) * stage_bytes + x_bytes);
        if (cnt == tile) {
            mbar_wait(&bars[s], parity);
        } else {  // ragged last tile: plain coalesced loads into the same staging buffers
            for (int j = tid; j < cnt * d; j += UPD_THREADS) xs[j] = ldg_stream(p.x + first * d + j);
            for (int j = tid; j < cnt; j += UPD_THREADS) ws[j] = ldg_stream(p.w_in + first + j);
            __syncthreads();
        }
        for (int j = tid; j < cnt; j += UPD_THREADS) {
            const double* xr = xs + static_cast<size_t>(j) * d;
            auto row = [&](int c) { return xr[c]; };
            auto meas = [&](int c) { return meas_s[c]; };
            const double L = model_likelihood<KIND, BINOM>(p.mv, p.ev, row, meas, lane);
            const double wn = ws[j] * inv_norm;  // previous step's normalisation, applied lazily
            const double wv = wn * L;            // smc.py:354
            stg_stream(p.w_out + first + j, wv);
            acc_s += wv;
            acc_q = fma(wv, wv, acc_q);
            acc_min = fmin(acc_min, wv);
            acc_bad += (wv >= 0.0) ? 0.0 : 1.0;  // counts negatives and NaNs (smc.py:416)
        }
        __syncthreads();  // every thread is done with stage s
        if (tid == 0 && i + UPD_STAGES < my_tiles) issue(i + UPD_STAGES);
    }

    block_reduce4(acc_s, acc_q, acc_min, acc_bad, red);
    if (tid == 0) {
        p.partials[blockIdx.x * 4 + 0] = acc_s;
        p.partials[blockIdx.x * 4 + 1] = acc_q;
        p.partials[blockIdx.x * 4 + 2] = acc_min;
        p.partials[blockIdx.x * 4 + 3] = acc_bad;
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        finish_stats(p.partials, gridDim.x, p.stats_out, red);
        if (tid == 0) *p.ticket = 0u;  // ready for the next launch on this stream
    }
}

// Tile size: ~16 KB per stage, a multiple of 16 particles so every bulk copy is 128-B granular.
static int choose_tile(int d) {
    int t = 2048 / (d + 1);
    t = (t / 16) * 16;
    if (t < 16) t = 16;
    if (t > 1024) t = 1024;
    return t;
}

static size_t update_smem_bytes(int d) {
    return 128 + QB_MAX_D * 8 + static_cast<size_t>(UPD_STAGES) * choose_tile(d) * (d + 1) * 8;
}

typedef void (*update_kernel_t)(const UpdateParams);

static update_kernel_t pick_update_kernel(const qb_model& m) {
    switch (m.kind) {
        case QB_MODEL_PRECESSION:
            return m.binomial ? fused_update_kernel<QB_MODEL_PRECESSION, true>
                              : fused_update_kernel<QB_MODEL_PRECESSION, false>;
        case QB_MODEL_RB:
            return m.binomial ? fused_update_kernel<QB_MODEL_RB, true> : fused_update_kernel<QB_MODEL_RB, false>;
        case QB_MODEL_TOMOGRAPHY:
            return m.binomial ? fused_update_kernel<QB_MODEL_TOMOGRAPHY, true>
                              : fused_update_kernel<QB_MODEL_TOMOGRAPHY, false>;
    }
    return nullptr;
}

int validate_model(const qb_model* m) {
    QB_REQUIRE(m != nullptr, QB_ERR_INVALID_ARGUMENT, "model is NULL");
    QB_REQUIRE(m->d >= 1 && m->d <= QB_MAX_D, QB_ERR_INVALID_ARGUMENT, "n_modelparams %d outside [1, %d]", m->d,
               QB_MAX_D);
    switch (m->kind) {
        case QB_MODEL_PRECESSION:
            QB_REQUIRE(m->d == 1, QB_ERR_UNSUPPORTED_MODEL, "precession model has 1 model parameter, got %d", m->d);
            break;
        case QB_MODEL_RB:
            QB_REQUIRE(m->d == (m->interleaved ? 4 : 3), QB_ERR_UNSUPPORTED_MODEL,
                       "RB model needs %d model parameters, got %d", m->interleaved ? 4 : 3, m->d);
            break;
        case QB_MODEL_TOMOGRAPHY:
            break;
        default:
            set_error("unknown model kind %d (no CPU fallback exists)", m->kind);
            return QB_ERR_UNSUPPORTED_MODEL;
    }
    return QB_OK;
}

struct UpdateLaunchCache {
    int blocks_per_sm[4][2];
    bool ready[4][2];
};
static UpdateLaunchCache g_cache = {};

static int update_grid_limit(const qb_model& m, update_kernel_t k, size_t smem) {
    const int bi = m.binomial ? 1 : 0;
    if (!g_cache.ready[m.kind][bi]) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
        g_cache.ready[m.kind][bi] = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, UPD_THREADS, smem) != cudaSuccess) return -1;
    if (per_sm < 1) per_sm = 1;
    return per_sm * sm_count();
}

}  // namespace qb

using namespace qb;

extern "C" size_t qb_update_workspace_bytes(int64_t n, int32_t d) {
    (void)n;
    (void)d;
    // per-block partials for up to 32 blocks/SM on up to 256 SMs + the ticket
    return static_cast<size_t>(32) * 256 * 4 * sizeof(double) + 256;
}

extern "C" int qb_fused_update(const qb_model* model, const qb_expparams* ep, int64_t outcome, const double* d_x,
                               int64_t n, const double* d_w_in, double* d_w_out, const double* d_stats_in,
                               double* d_stats_out, void* d_ws, size_t ws_bytes, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(ep && d_x && d_w_in && d_w_out && d_stats_in && d_stats_out && d_ws, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: NULL pointer argument");
    QB_REQUIRE(n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: n must be >= 1, got %lld", (long long)n);
    QB_REQUIRE(ws_bytes >= qb_update_workspace_bytes(n, model->d), QB_ERR_WORKSPACE,
               "qb_fused_update: workspace too small");
    QB_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w_in) & 15) == 0,
               QB_ERR_INVALID_ARGUMENT, "qb_fused_update: x and w must be 16-byte aligned");

    UpdateParams p;
    p.x = d_x;
    p.w_in = d_w_in;
    p.w_out = d_w_out;
    p.stats_in = d_stats_in;
    p.stats_out = d_stats_out;
    p.ticket = reinterpret_cast<unsigned int*>(d_ws);
    p.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    p.n = n;
    p.d = model->d;
    p.tile = choose_tile(model->d);
    p.mv = make_model_view(*model);
    p.ev = make_exp_view(*model, *ep, outcome);
    for (int c = 0; c < QB_MAX_D; ++c) p.meas[c] = (c < model->d) ? ep->meas[c] : 0.0;

    update_kernel_t k = pick_update_kernel(*model);
    const size_t smem = update_smem_bytes(model->d);
    const int limit = update_grid_limit(*model, k, smem);
    QB_REQUIRE(limit > 0, QB_ERR_CUDA, "qb_fused_update: occupancy query failed: %s",
               cudaGetErrorString(cudaGetLastError()));
    const int64_t ntiles = (n + p.tile - 1) / p.tile;
    int grid = static_cast<int>(ntiles < limit ? ntiles : limit);
    if (grid > 32 * 256) grid = 32 * 256;
    k<<<grid, UPD_THREADS, smem, as_stream(stream)>>>(p);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
