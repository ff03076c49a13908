// Time-dependent and noisy model decorators:
//   qb_walk_step          Model.update_timestep of RandomWalkModel / GaussianRandomWalkModel
//                         (derived_models.py:733-741, 921-963) and DiffusiveTomographyModel (tomography/models.py:
//                         257-272, the canonicalisation that follows is qb_tomo_canonicalize_ld), in place
//   qb_poison_likelihood  PoisonedModel.likelihood's clipped Gaussian noise (derived_models.py:188-204)
// Both are element-wise passes over data the fused update has just touched; the normals come from a buffer so that
// the caller chooses the generator (host np.random stream for parity with the reference, device MT19937 or Philox).
//
// Compiled with --fmad=false: one rounding per reference ufunc.
#include "qb_common.cuh"

namespace qb {

constexpr int WALK_MAX = 16;   // walking parameters per launch (larger sets are split by the host wrapper)

struct WalkParams {
    double* x;
    const double* z;
    int64_t n;
    int32_t d, n_rw, mode, kz;
    double pre, mult;
    int32_t idx[WALK_MAX], zcol[WALK_MAX], sidx[WALK_MAX];
    double scale[WALK_MAX];
};

__global__ void __launch_bounds__(256) walk_step_kernel(const __grid_constant__ WalkParams p) {
    const int64_t total = p.n * p.n_rw;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / p.n_rw;
        const int c = static_cast<int>(e - i * p.n_rw);
        const double zv = p.z[i * p.kz + p.zcol[c]];
        double step;
        if (p.mode == QB_WALK_ADD)
            step = zv;                                                        // the model's own step distribution
        else if (p.mode == QB_WALK_FIXED)
            step = p.scale[c] * zv;                                           // derived_models.py:925-929
        else
            step = (p.x[i * p.d + p.sidx[c]] * p.pre) * zv;                   // :926 / tomography/models.py:261-264
        step = p.mult * step;                                                 // derived_models.py:944-945
        double* dst = p.x + i * p.d + p.idx[c];
        *dst = *dst + step;                                                   // :961-962 / tomography/models.py:267
    }
}

__global__ void __launch_bounds__(256) poison_kernel(double* __restrict__ L, int64_t n, const double* __restrict__ z,
                                                     int mode, double tol, double denom) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double l = L[i];
        // epsilon *= tol   |   epsilon *= sqrt(p * (1 - p) / (N + 2 hedge + 1))   (utils.py:683-688)
        const double sigma = (mode == 0) ? tol : sqrt((l * (1.0 - l)) / denom);
        double v = l + z[i] * sigma;
        v = (v < 0.0) ? 0.0 : ((v > 1.0) ? 1.0 : v);                          // np.clip(L + epsilon, 0, 1)
        L[i] = v;
    }
}

static int walk_grid(int64_t work) {
    int64_t want = (work + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (want > cap) want = cap;
    return static_cast<int>(want < 1 ? 1 : want);
}

}  // namespace qb

using namespace qb;

extern "C" int qb_walk_step(double* d_x, int64_t n, int32_t d, int32_t n_rw, const int32_t* h_idx,
                            const int32_t* h_zcol, int32_t mode, const double* h_scale, const int32_t* h_sidx,
                            double pre, double mult, const double* d_z, int32_t kz, void* stream) {
    QB_REQUIRE(d_x && d_z && h_idx && h_zcol && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_walk_step: bad arguments");
    QB_REQUIRE(d >= 1 && d <= QB_MAX_D && n_rw >= 1 && n_rw <= d && kz >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_walk_step: bad sizes (d %d, n_rw %d, kz %d)", d, n_rw, kz);
    QB_REQUIRE(mode == QB_WALK_ADD || mode == QB_WALK_FIXED || mode == QB_WALK_LEARNED, QB_ERR_INVALID_ARGUMENT,
               "qb_walk_step: unknown mode %d", mode);
    QB_REQUIRE(mode != QB_WALK_FIXED || h_scale, QB_ERR_INVALID_ARGUMENT, "qb_walk_step: NULL scales");
    QB_REQUIRE(mode != QB_WALK_LEARNED || h_sidx, QB_ERR_INVALID_ARGUMENT, "qb_walk_step: NULL scale columns");
    for (int c0 = 0; c0 < n_rw; c0 += WALK_MAX) {
        WalkParams p;
        p.x = d_x;
        p.z = d_z;
        p.n = n;
        p.d = d;
        p.n_rw = (n_rw - c0 < WALK_MAX) ? n_rw - c0 : WALK_MAX;
        p.mode = mode;
        p.kz = kz;
        p.pre = pre;
        p.mult = mult;
        for (int c = 0; c < WALK_MAX; ++c) {
            const bool in = c < p.n_rw;
            p.idx[c] = in ? h_idx[c0 + c] : 0;
            p.zcol[c] = in ? h_zcol[c0 + c] : 0;
            p.sidx[c] = (in && h_sidx) ? h_sidx[c0 + c] : 0;
            p.scale[c] = (in && h_scale) ? h_scale[c0 + c] : 0.0;
            QB_REQUIRE(!in || (p.idx[c] >= 0 && p.idx[c] < d && p.zcol[c] >= 0 && p.zcol[c] < kz && p.sidx[c] >= 0 &&
                               p.sidx[c] < d),
                       QB_ERR_INVALID_ARGUMENT, "qb_walk_step: index out of range");
        }
        walk_step_kernel<<<walk_grid(n * p.n_rw), 256, 0, as_stream(stream)>>>(p);
        QB_CUDA_CHECK(cudaGetLastError());
    }
    return QB_OK;
}

extern "C" int qb_poison_likelihood(double* d_L, int64_t n, const double* d_z, int32_t mode, double tol, double denom,
                                    void* stream) {
    QB_REQUIRE(d_L && d_z && n >= 1 && (mode == 0 || mode == 1), QB_ERR_INVALID_ARGUMENT,
               "qb_poison_likelihood: bad arguments");
    QB_REQUIRE(mode == 0 || denom > 0.0, QB_ERR_INVALID_ARGUMENT, "qb_poison_likelihood: denom must be positive");
    poison_kernel<<<walk_grid(n), 256, 0, as_stream(stream)>>>(d_L, n, d_z, mode, tol, denom);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
