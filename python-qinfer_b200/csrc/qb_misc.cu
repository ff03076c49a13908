// Library plumbing, weight utilities, plain likelihood and validity kernels.
#include <cstdarg>
#include <cstring>
#include "qb_models.cuh"

namespace qb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return QB_ERR_CUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached[dev] = n;
    }
    return cached[dev];
}

int validate_model(const qb_model* m);  // qb_update.cu

// ---- weights ------------------------------------------------------------------
__global__ void set_uniform_kernel(double* w, int64_t n, int64_t n_global, double* stats) {
    const double v = 1.0 / static_cast<double>(n_global);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = v;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[QB_STAT_NORM] = 1.0;
        stats[QB_STAT_SUMSQ] = v;
        stats[QB_STAT_MIN] = v;
        stats[QB_STAT_NBAD] = 0.0;
        stats[QB_STAT_INV_NORM] = 1.0;
        stats[QB_STAT_NESS] = static_cast<double>(n_global);
        stats[QB_STAT_TAG] = 0.0;
        stats[QB_STAT_SKIPPED] = 0.0;
        stats[QB_STAT_ATTN] = 0.0;
    }
}

__global__ void normalized_kernel(const double* __restrict__ w, int64_t n, const double* __restrict__ stats,
                                  double* __restrict__ out) {
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = w[i] * inv;
}

// restat / clip share one reduction kernel: CLIP also rewrites w in place.
template <bool CLIP>
__global__ void __launch_bounds__(256) restat_kernel(double* w, int64_t n, const double* stats_in, double* partials) {
    __shared__ double red[8 * 4];
    const double inv = CLIP ? stats_in[QB_STAT_INV_NORM] : 1.0;
    double s = 0.0, q = 0.0, mn = INFINITY, bad = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = w[i];
        if (CLIP) {
            v = v * inv;
            // np.clip(weights, 0, 1) (smc.py:418); NaN propagates like NumPy's clip
            v = (v < 0.0) ? 0.0 : ((v > 1.0) ? 1.0 : v);
            w[i] = v;
        }
        s += v;
        q = fma(v, v, q);
        mn = fmin(mn, v);
        bad += (v >= 0.0) ? 0.0 : 1.0;
    }
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
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            s += red[k * 4 + 0];
            q += red[k * 4 + 1];
            mn = fmin(mn, red[k * 4 + 2]);
            bad += red[k * 4 + 3];
        }
        partials[blockIdx.x * 4 + 0] = s;
        partials[blockIdx.x * 4 + 1] = q;
        partials[blockIdx.x * 4 + 2] = mn;
        partials[blockIdx.x * 4 + 3] = bad;
    }
}

__global__ void restat_finish_kernel(const double* partials, int nblocks, double* stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0, q = 0.0, mn = INFINITY, bad = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        s += partials[b * 4 + 0];
        q += partials[b * 4 + 1];
        mn = fmin(mn, partials[b * 4 + 2]);
        bad += partials[b * 4 + 3];
    }
    stats[QB_STAT_NORM] = s;
    stats[QB_STAT_SUMSQ] = q;
    stats[QB_STAT_MIN] = mn;
    stats[QB_STAT_NBAD] = bad;
    stats[QB_STAT_INV_NORM] = 1.0;  // weights are taken as given
    stats[QB_STAT_NESS] = 1.0 / q;  // distributions.py:299-307 on the weights as they are
    stats[QB_STAT_TAG] = 0.0;
    stats[QB_STAT_SKIPPED] = 0.0;
    stats[QB_STAT_ATTN] = 0.0;
}

// ---- plain likelihood + validity -----------------------------------------------
struct LikParams {
    const double* x;
    double* out;   // L + (o * n) * n_e + e ; element i at stride n_e
    int64_t n;
    int32_t n_e;
    ModelView mv;
    ExpView ev;
    double meas[QB_MAX_D];
};

template <int KIND, bool BINOM>
__global__ void __launch_bounds__(256) likelihood_kernel(const __grid_constant__ LikParams p) {
    const int d = p.mv.d;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double* xr = p.x + i * d;
        auto row = [&](int c) { return xr[c]; };
        auto meas = [&](int c) { return p.meas[c]; };
        p.out[i * p.n_e] = model_likelihood<KIND, BINOM>(p.mv, p.ev, row, meas, 0);
    }
}

__global__ void __launch_bounds__(256) valid_kernel(ModelView mv, const double* __restrict__ x, int64_t n,
                                                    uint8_t* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* xr = x + i * mv.d;
        auto row = [&](int c) { return xr[c]; };
        out[i] = model_valid(mv, row) ? 1 : 0;
    }
}

// ---- hypothetical update (smc.py:324-386) for one (outcome, experiment) pair ---------------------------
struct HypParams {
    const double* x;
    const double* w;
    const double* stats;
    double* out;        // hyp weights for this pair, n contiguous doubles
    double* L;          // likelihoods for this pair (may be NULL)
    double* partials;   // [grid]
    int64_t n;
    ModelView mv;
    ExpView ev;
    double meas[QB_MAX_D];
};

template <int KIND, bool BINOM>
__global__ void __launch_bounds__(256) hyp_kernel(const __grid_constant__ HypParams p) {
    __shared__ double red[8];
    const int d = p.mv.d;
    const double inv = p.stats[QB_STAT_INV_NORM];
    double s = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double* xr = p.x + i * d;
        auto row = [&](int c) { return xr[c]; };
        auto meas = [&](int c) { return p.meas[c]; };
        const double L = model_likelihood<KIND, BINOM>(p.mv, p.ev, row, meas, 0);
        const double h = (p.w[i] * inv) * L;  // weights * L (smc.py:354)
        p.out[i] = h;
        if (p.L != nullptr) p.L[i] = L;
        s += h;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) s += red[k];
        p.partials[blockIdx.x] = s;
    }
}

__global__ void hyp_finish_kernel(const double* partials, int nblocks, double* norm_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[b];
    *norm_out = s;  // norm_scale (smc.py:357)
}

__global__ void __launch_bounds__(256) hyp_scale_kernel(double* out, int64_t n, const double* norm) {
    const double eps = 2.220446049250313e-16;
    const double nv = *norm;
    const double div = (fabs(nv) < eps) ? 1.0 : nv;  // fixed_norm_scale (smc.py:369-370)
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = out[i] / div;  // smc.py:373
}

static int grid_for(int64_t n, int threads, int per_sm) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace qb

using namespace qb;

extern "C" int qb_abi_version(void) { return QB_ABI_VERSION; }
extern "C" const char* qb_last_error(void) { return qb::g_err; }
extern "C" int qb_device_sm_count(void) { return qb::sm_count(); }
extern "C" void qb_struct_sizes(int32_t out[3]) {
    out[0] = static_cast<int32_t>(sizeof(qb_model));
    out[1] = static_cast<int32_t>(sizeof(qb_expparams));
    out[2] = static_cast<int32_t>(sizeof(qb_update_ctl));
}

extern "C" int qb_weights_set_uniform(double* d_w, int64_t n, double* d_stats, void* stream) {
    return qb_weights_set_uniform_global(d_w, n, n, d_stats, stream);
}

extern "C" int qb_weights_set_uniform_global(double* d_w, int64_t n_local, int64_t n_global, double* d_stats,
                                             void* stream) {
    QB_REQUIRE(d_w && d_stats && n_local >= 1 && n_global >= n_local, QB_ERR_INVALID_ARGUMENT,
               "qb_weights_set_uniform: bad arguments");
    set_uniform_kernel<<<grid_for(n_local, 256, 8), 256, 0, as_stream(stream)>>>(d_w, n_local, n_global, d_stats);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_weights_normalized(const double* d_w, int64_t n, const double* d_stats, double* d_out,
                                     void* stream) {
    QB_REQUIRE(d_w && d_stats && d_out && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_weights_normalized: bad arguments");
    normalized_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(d_w, n, d_stats, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

static int restat_common(double* d_w, int64_t n, double* d_stats, double* d_ws, size_t ws_bytes, void* stream,
                         bool clip) {
    QB_REQUIRE(d_w && d_stats && d_ws && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_weights_restat/clip: bad arguments");
    const int grid = grid_for(n, 256, 8);
    QB_REQUIRE(ws_bytes >= static_cast<size_t>(grid) * 4 * sizeof(double) + 256, QB_ERR_WORKSPACE,
               "qb_weights_restat/clip: workspace too small");
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    if (clip)
        restat_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(d_w, n, d_stats, partials);
    else
        restat_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(d_w, n, d_stats, partials);
    QB_CUDA_CHECK(cudaGetLastError());
    restat_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(partials, grid, d_stats);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_weights_restat(const double* d_w, int64_t n, double* d_stats, double* d_ws, size_t ws_bytes,
                                 void* stream) {
    return restat_common(const_cast<double*>(d_w), n, d_stats, d_ws, ws_bytes, stream, false);
}

extern "C" int qb_weights_clip(double* d_w, int64_t n, double* d_stats, double* d_ws, size_t ws_bytes, void* stream) {
    return restat_common(d_w, n, d_stats, d_ws, ws_bytes, stream, true);
}

extern "C" int qb_likelihood(const qb_model* model, const qb_expparams* eps, int32_t n_e, const int64_t* outcomes,
                             int32_t n_o, const double* d_x, int64_t n, double* d_L, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(eps && outcomes && d_x && d_L && n >= 1 && n_e >= 1 && n_o >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_likelihood: bad arguments");
    LikParams p;
    p.x = d_x;
    p.n = n;
    p.n_e = n_e;
    p.mv = make_model_view(*model);
    const int grid = grid_for(n, 256, 8);
    for (int o = 0; o < n_o; ++o) {
        for (int e = 0; e < n_e; ++e) {
            p.ev = make_exp_view(*model, eps[e], outcomes[o]);
            for (int c = 0; c < QB_MAX_D; ++c) p.meas[c] = (c < model->d) ? eps[e].meas[c] : 0.0;
            p.out = d_L + static_cast<int64_t>(o) * n * n_e + e;
#define QB_LAUNCH_LIK(K)                                                                        \
    if (model->binomial)                                                                        \
        likelihood_kernel<K, true><<<grid, 256, 0, as_stream(stream)>>>(p);                     \
    else                                                                                        \
        likelihood_kernel<K, false><<<grid, 256, 0, as_stream(stream)>>>(p);
            if (model->kind == QB_MODEL_PRECESSION) {
                QB_LAUNCH_LIK(QB_MODEL_PRECESSION)
            } else if (model->kind == QB_MODEL_RB) {
                QB_LAUNCH_LIK(QB_MODEL_RB)
            } else if (model->kind == QB_MODEL_COIN) {
                QB_LAUNCH_LIK(QB_MODEL_COIN)
            } else {
                QB_LAUNCH_LIK(QB_MODEL_TOMOGRAPHY)
            }
#undef QB_LAUNCH_LIK
            QB_CUDA_CHECK(cudaGetLastError());
        }
    }
    return QB_OK;
}

extern "C" int qb_hypothetical_update(const qb_model* model, const qb_expparams* eps, int32_t n_e,
                                      const int64_t* outcomes, int32_t n_o, const double* d_x, const double* d_w,
                                      const double* d_stats, int64_t n, double* d_weights, double* d_L,
                                      double* d_norms, void* d_ws, size_t ws_bytes, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(eps && outcomes && d_x && d_w && d_stats && d_weights && d_norms && d_ws && n >= 1 && n_e >= 1 &&
                   n_o >= 1,
               QB_ERR_INVALID_ARGUMENT, "qb_hypothetical_update: bad arguments");
    const int grid = grid_for(n, 256, 8);
    QB_REQUIRE(ws_bytes >= static_cast<size_t>(grid) * sizeof(double) + 256, QB_ERR_WORKSPACE,
               "qb_hypothetical_update: workspace too small");
    HypParams p;
    p.x = d_x;
    p.w = d_w;
    p.stats = d_stats;
    p.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    p.n = n;
    p.mv = make_model_view(*model);
    cudaStream_t st = as_stream(stream);
    for (int o = 0; o < n_o; ++o) {
        for (int e = 0; e < n_e; ++e) {
            const int64_t pair = static_cast<int64_t>(o) * n_e + e;  // output layout (n_outcomes, n_expparams, n)
            p.ev = make_exp_view(*model, eps[e], outcomes[o]);
            for (int c = 0; c < QB_MAX_D; ++c) p.meas[c] = (c < model->d) ? eps[e].meas[c] : 0.0;
            p.out = d_weights + pair * n;
            p.L = d_L ? d_L + pair * n : nullptr;
#define QB_LAUNCH_HYP(K)                                                    \
    if (model->binomial)                                                    \
        hyp_kernel<K, true><<<grid, 256, 0, st>>>(p);                       \
    else                                                                    \
        hyp_kernel<K, false><<<grid, 256, 0, st>>>(p);
            if (model->kind == QB_MODEL_PRECESSION) {
                QB_LAUNCH_HYP(QB_MODEL_PRECESSION)
            } else if (model->kind == QB_MODEL_RB) {
                QB_LAUNCH_HYP(QB_MODEL_RB)
            } else if (model->kind == QB_MODEL_COIN) {
                QB_LAUNCH_HYP(QB_MODEL_COIN)
            } else {
                QB_LAUNCH_HYP(QB_MODEL_TOMOGRAPHY)
            }
#undef QB_LAUNCH_HYP
            QB_CUDA_CHECK(cudaGetLastError());
            hyp_finish_kernel<<<1, 32, 0, st>>>(p.partials, grid, d_norms + pair);
            hyp_scale_kernel<<<grid, 256, 0, st>>>(p.out, n, d_norms + pair);
            QB_CUDA_CHECK(cudaGetLastError());
        }
    }
    return QB_OK;
}

extern "C" int qb_are_models_valid(const qb_model* model, const double* d_x, int64_t n, uint8_t* d_valid,
                                   void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x && d_valid && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_are_models_valid: bad arguments");
    valid_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(make_model_view(*model), d_x, n, d_valid);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
