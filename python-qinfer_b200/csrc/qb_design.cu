// Experiment-design reductions over the particle cloud: the Bayes risk and the expected
// information gain of hypothetical experiments (smc.py:553-605, 607-657), computed where the cloud lives.
//
// For ONE experiment and its outcome list os[0..n_o) the reference forms, for every outcome o,
//   L_o,i   = Model.likelihood(os[o], x_i)            for o < n_o - 1
//   L_last,i = 1 - sum_{o < n_o - 1} L_o,i             (smc.py:589: "the likelihood over outcomes sums to 1")
//   N_o     = sum_i w_i L_o,i ;  w_hyp_o,i = w_i L_o,i / N_o
// Here the last outcome's likelihood is EVALUATED like the others: for the two-outcome models that is the same
// complement (pr1 = 1 - pr0); for BinomialModel it replaces a difference that cancels to +-1e-16 (and then feeds
// log() a zero or negative number: the reference returns NaN there) by the pmf itself.  Under MLEModel (a likelihood
// power != 1) the likelihoods do not sum to one, so there the reference's complement is formed as written.
// and then either the posterior variance of every outcome (risk) or its KL divergence from the prior (gain).
// The (n_o, n) tensors w_hyp and L never exist here: pass 1 reduces, per outcome,
//   A = sum h,  B_j = sum h (x_j - c_j),  C_j = sum h (x_j - c_j)^2        with h = w_i L_o,i
// around a centre c (the current posterior mean, so that var = C/A - (B/A)^2 does not cancel), and pass 2, given
// N_o = A, reduces  K = sum w_hyp log(w_hyp / w)  elementwise exactly as smc.py:651 writes it.
// grid.y = outcome: every block row streams the cloud once.  Compiled with --fmad=false like every model kernel.
#include "qb_models.cuh"

namespace qb {

constexpr int DSN_THREADS = 256;

struct DesignParams {
    const double* x;
    const double* w;
    const double* stats;
    const ExpView* evs;     // device: n_o views (one per outcome) of this experiment
    double* partials;       // [n_o][grid.x][nv]
    const double* sums;     // pass 2: the finished pass-1 sums, [n_o][1 + 2 d]
    int64_t n;
    int32_t n_o, d, nv, pass;
    ModelView mv;
    double meas[QB_MAX_D];
    double centre[QB_MAX_D];
};

template <int KIND, bool BINOM>
__device__ __forceinline__ double outcome_likelihood(const DesignParams& p, const double* xr, int o) {
    auto row = [&](int c) { return xr[c]; };
    auto meas = [&](int c) { return p.meas[c]; };
    if (p.mv.like_pow != 1.0 && o == p.n_o - 1 && p.n_o > 1) {
        // MLEModel: the powered likelihoods no longer sum to one, and the reference DEFINES the last outcome's as
        // 1 - sum of the others (smc.py:589) — sin^4 versus 1 - cos^4: evaluating it would silently change the risk
        double s = 0.0;
        for (int q = 0; q < p.n_o - 1; ++q) s += model_likelihood<KIND, BINOM>(p.mv, p.evs[q], row, meas, 0);
        return 1.0 - s;
    }
    return model_likelihood<KIND, BINOM>(p.mv, p.evs[o], row, meas, 0);
}

// DMAX: compile-time bound on d for the per-thread accumulators (1, 4, 16, 64).
template <int KIND, bool BINOM, int DMAX>
__global__ void __launch_bounds__(DSN_THREADS) design_sums_kernel(const __grid_constant__ DesignParams p) {
    __shared__ double red[DSN_THREADS / 32];
    const int o = blockIdx.y;
    const int d = p.d;
    const double inv = p.stats[QB_STAT_INV_NORM];
    double A = 0.0, B[DMAX], C[DMAX];
#pragma unroll
    for (int c = 0; c < DMAX; ++c) B[c] = C[c] = 0.0;
    double norm = 1.0, div = 1.0;
    if (p.pass == 2) {
        norm = p.sums[static_cast<size_t>(o) * (1 + 2 * d)];
        // hypothetical_update divides by the guarded normalisation (smc.py:369-373) for the outcomes it was given;
        // the complement outcome is divided by its plain sum (smc.py:591)
        div = (o < p.n_o - 1 && fabs(norm) < 2.220446049250313e-16) ? 1.0 : norm;
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double* xr = p.x + i * d;
        const double wn = p.w[i] * inv;
        const double L = outcome_likelihood<KIND, BINOM>(p, xr, o);
        const double h = wn * L;  // smc.py:354
        if (p.pass == 1) {
            A += h;
#pragma unroll
            for (int c = 0; c < DMAX; ++c) {
                if (c < d) {
                    const double dx = xr[c] - p.centre[c];
                    const double t = h * dx;
                    B[c] += t;
                    C[c] = fma(t, dx, C[c]);
                }
            }
        } else {
            const double wh = h / div;
            if (wh > 0.0) A += wh * log(wh / wn);  // smc.py:651 term by term; a zero-likelihood particle adds its limit 0
        }
    }
    const int nv = (p.pass == 1) ? 1 + 2 * d : 1;
    double* out = p.partials + (static_cast<size_t>(o) * gridDim.x + blockIdx.x) * p.nv;
    for (int v = 0; v < nv; ++v) {
        double t;
        if (v == 0) {
            t = A;
        } else {
            // static indexing keeps B and C in registers
            t = 0.0;
#pragma unroll
            for (int c = 0; c < DMAX; ++c) {
                if (c < d && v == 1 + c) t = B[c];
                if (c < d && v == 1 + d + c) t = C[c];
            }
        }
        t = warp_sum(t);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < DSN_THREADS / 32; ++k) t += red[k];
            out[v] = t;
        }
        __syncthreads();
    }
}

// out[o][v] = sum over blocks in block order (deterministic); one warp per (o, v)
__global__ void __launch_bounds__(32) design_finish_kernel(const double* __restrict__ partials, int nblocks, int nv_stride,
                                                           int nv, double* __restrict__ out, int out_stride) {
    const int o = blockIdx.y, v = blockIdx.x;
    if (v >= nv) return;
    double t = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32)
        t += partials[(static_cast<size_t>(o) * nblocks + b) * nv_stride + v];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[static_cast<size_t>(o) * out_stride + v] = t;
}

int validate_model(const qb_model* m);

static int design_grid(int64_t n, int n_o) {
    int64_t g = (n + DSN_THREADS - 1) / DSN_THREADS;
    int64_t cap = (static_cast<int64_t>(sm_count()) * 8 + n_o - 1) / n_o;
    if (cap < 4) cap = 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

typedef void (*design_kernel_t)(const DesignParams);

template <int KIND>
static design_kernel_t pick_design_k(bool binom, int d) {
#define QB_PICK(DM) (binom ? design_sums_kernel<KIND, true, DM> : design_sums_kernel<KIND, false, DM>)
    if (d <= 1) return QB_PICK(1);
    if (d <= 4) return QB_PICK(4);
    if (d <= 16) return QB_PICK(16);
    return QB_PICK(64);
#undef QB_PICK
}

static design_kernel_t pick_design(const qb_model& m) {
    switch (m.kind) {
        case QB_MODEL_PRECESSION: return pick_design_k<QB_MODEL_PRECESSION>(m.binomial != 0, m.d);
        case QB_MODEL_RB: return pick_design_k<QB_MODEL_RB>(m.binomial != 0, m.d);
        case QB_MODEL_TOMOGRAPHY: return pick_design_k<QB_MODEL_TOMOGRAPHY>(m.binomial != 0, m.d);
        case QB_MODEL_COIN: return pick_design_k<QB_MODEL_COIN>(m.binomial != 0, m.d);
    }
    return nullptr;
}

}  // namespace qb

using namespace qb;

extern "C" size_t qb_design_workspace_bytes(int64_t n, int32_t d, int32_t n_o) {
    if (n < 1 || d < 1 || n_o < 1) return 0;
    const size_t evs = ((static_cast<size_t>(n_o) * sizeof(ExpView) + 255) / 256) * 256;
    const size_t part = static_cast<size_t>(n_o) * design_grid(n, n_o) * (1 + 2 * d) * sizeof(double);
    return 256 + evs + part + 256;
}

extern "C" int qb_design_sums(const qb_model* model, const qb_expparams* ep, const int64_t* outcomes, int32_t n_o,
                              const double* d_x, const double* d_w, const double* d_stats, int64_t n,
                              const double* h_centre, double* d_sums, double* d_kld, void* d_ws, size_t ws_bytes,
                              void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(ep && outcomes && d_x && d_w && d_stats && h_centre && d_sums && d_ws && n >= 1 && n_o >= 1,
               QB_ERR_INVALID_ARGUMENT, "qb_design_sums: bad arguments");
    QB_REQUIRE(n_o <= 65535, QB_ERR_INVALID_ARGUMENT, "qb_design_sums: at most 65535 outcomes, got %d", n_o);
    QB_REQUIRE(ws_bytes >= qb_design_workspace_bytes(n, model->d, n_o), QB_ERR_WORKSPACE,
               "qb_design_sums: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int d = model->d;
    unsigned char* base = reinterpret_cast<unsigned char*>(d_ws) + 256;
    const size_t evs_bytes = ((static_cast<size_t>(n_o) * sizeof(ExpView) + 255) / 256) * 256;
    // the outcome views travel through a pageable staging vector: cudaMemcpyAsync copies it out before returning
    static thread_local ExpView* h_evs = nullptr;
    static thread_local int h_cap = 0;
    if (h_cap < n_o) {
        delete[] h_evs;
        h_evs = new ExpView[n_o];
        h_cap = n_o;
    }
    for (int o = 0; o < n_o; ++o) h_evs[o] = make_exp_view(*model, *ep, outcomes[o]);
    QB_CUDA_CHECK(cudaMemcpyAsync(base, h_evs, static_cast<size_t>(n_o) * sizeof(ExpView), cudaMemcpyHostToDevice, st));

    DesignParams p;
    p.x = d_x;
    p.w = d_w;
    p.stats = d_stats;
    p.evs = reinterpret_cast<const ExpView*>(base);
    p.partials = reinterpret_cast<double*>(base + evs_bytes);
    p.sums = d_sums;
    p.n = n;
    p.n_o = n_o;
    p.d = d;
    p.nv = 1 + 2 * d;
    p.pass = 1;
    p.mv = make_model_view(*model);
    for (int c = 0; c < QB_MAX_D; ++c) {
        p.meas[c] = (c < d) ? ep->meas[c] : 0.0;
        p.centre[c] = (c < d) ? h_centre[c] : 0.0;
    }
    design_kernel_t k = pick_design(*model);
    QB_REQUIRE(k != nullptr, QB_ERR_UNSUPPORTED_MODEL, "qb_design_sums: unsupported model kind %d", model->kind);
    const int gx = design_grid(n, n_o);
    const dim3 grid(gx, n_o);
    k<<<grid, DSN_THREADS, 0, st>>>(p);
    QB_CUDA_CHECK(cudaGetLastError());
    design_finish_kernel<<<dim3(p.nv, n_o), 32, 0, st>>>(p.partials, gx, p.nv, p.nv, d_sums, p.nv);
    QB_CUDA_CHECK(cudaGetLastError());
    if (d_kld != nullptr) {
        p.pass = 2;
        k<<<grid, DSN_THREADS, 0, st>>>(p);
        QB_CUDA_CHECK(cudaGetLastError());
        design_finish_kernel<<<dim3(1, n_o), 32, 0, st>>>(p.partials, gx, p.nv, 1, d_kld, 1);
        QB_CUDA_CHECK(cudaGetLastError());
    }
    return QB_OK;
}
