// Weighted first and second moments of the particle cloud in one pass
// (distributions.py:337-348, 386-387):
//   out[0] = sum_i w_i, out[1+m] = sum_i w_i x_im, out[1+d+m*d+n] = sum_i w_i x_im x_in
// with w_i the normalised weight.  HBM traffic 8(d+1) B/particle.
//
// Three kernels:
//   d <= 4   : per-thread register accumulators, streaming coalesced loads.
//   d == 16  : the d x d contraction runs on the FP64 tensor cores (DMMA,
//              mma.sync.m8n8k4.f64 — tcgen05 has no f64 kind): A = (w o X)^T tiles,
//              B = X tiles, three 8x8 accumulator blocks per warp (symmetry);
//              fragments are loaded straight from global memory, each warp load
//              covering 4 rows x 64 contiguous bytes (8 fully used sectors).
//   other d  : shared-memory staged generic kernel (one thread per output entry).
// Per-block partials are reduced in fixed block order by a finishing kernel, so
// results are deterministic for a given launch configuration.
#include "qb_common.cuh"

namespace qb {

constexpr int MOM_THREADS = 256;

// ---- small d ---------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(MOM_THREADS) moments_small_kernel(const double* __restrict__ x,
                                                                    const double* __restrict__ w,
                                                                    const double* __restrict__ stats, int64_t n,
                                                                    double* __restrict__ partials) {
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;
    __shared__ double red[(MOM_THREADS / 32) * NOUT];
    const double inv = stats[QB_STAT_INV_NORM];
    double acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double wi = ldg_stream(w + i) * inv;
        double xr[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xr[c] = ldg_stream(x + i * D + c);
        acc[0] += wi;
        int o = 1 + D;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const double wx = wi * xr[m];
            acc[1 + m] += wx;
#pragma unroll
            for (int c = m; c < D; ++c) {
                acc[o] = fma(wx, xr[c], acc[o]);
                ++o;
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[wid * NOUT + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NOUT) {
        double v = 0.0;
        for (int k = 0; k < MOM_THREADS / 32; ++k) v += red[k * NOUT + threadIdx.x];
        partials[static_cast<int64_t>(blockIdx.x) * NOUT + threadIdx.x] = v;
    }
}

// ---- d == 16: DMMA ------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Packed upper-triangular index helper (row-major packing, m <= c).
__host__ __device__ inline int tri_index(int m, int c, int d) { return m * d - (m * (m - 1)) / 2 + (c - m); }

__global__ void __launch_bounds__(MOM_THREADS) moments_d16_kernel(const double* __restrict__ x,
                                                                  const double* __restrict__ w,
                                                                  const double* __restrict__ stats, int64_t n,
                                                                  double* __restrict__ partials) {
    constexpr int D = 16;
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;  // 153
    constexpr int NW = MOM_THREADS / 32;
    __shared__ double cs[NW][D][D + 1];  // per-warp full second-moment block
    __shared__ double ms[NW][D + 1];     // per-warp sum w x (16) and sum w
    const double inv = stats[QB_STAT_INV_NORM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int kq = lane & 3;   // which of the 4 particles of a k-step this lane loads
    const int mr = lane >> 2;  // which of the 8 rows/cols of the fragment
    double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c11a = 0, c11b = 0;
    double mlo = 0, mhi = 0, sw = 0;
    const int64_t nsteps = (n + 3) / 4;
    const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * NW + wid;
    const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * NW;
    // UNROLL k-steps per iteration with all their loads issued first: one step keeps only 512 B per warp in
    // flight, far too little to cover HBM latency (r1: 89 us -> the loads, not the DMMAs, were the limit).
    constexpr int UNROLL = 4;
    for (int64_t s0 = warp_global * UNROLL; s0 < nsteps; s0 += warp_stride * UNROLL) {
        double wi[UNROLL], xlo[UNROLL], xhi[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = (s0 + u) * 4 + kq;
            wi[u] = 0.0;
            xlo[u] = 0.0;
            xhi[u] = 0.0;
            if (i < n) {
                wi[u] = ldg_stream(w + i);
                xlo[u] = ldg_stream(x + i * D + mr);
                xhi[u] = ldg_stream(x + i * D + 8 + mr);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const double wn = wi[u] * inv;
            const double alo = wn * xlo[u], ahi = wn * xhi[u];
            dmma_m8n8k4(c00a, c00b, alo, xlo[u]);
            dmma_m8n8k4(c01a, c01b, alo, xhi[u]);
            dmma_m8n8k4(c11a, c11b, ahi, xhi[u]);
            mlo += alo;
            mhi += ahi;
            sw += wn;
        }
    }
    // accumulator fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
    const int r = lane >> 2, c = (lane & 3) * 2;
    cs[wid][r][c] = c00a;
    cs[wid][r][c + 1] = c00b;
    cs[wid][r][8 + c] = c01a;
    cs[wid][r][8 + c + 1] = c01b;
    cs[wid][8 + r][8 + c] = c11a;
    cs[wid][8 + r][8 + c + 1] = c11b;
    // means: reduce over the 4 lanes that share mr
    mlo += __shfl_xor_sync(0xffffffffu, mlo, 1);
    mlo += __shfl_xor_sync(0xffffffffu, mlo, 2);
    mhi += __shfl_xor_sync(0xffffffffu, mhi, 1);
    mhi += __shfl_xor_sync(0xffffffffu, mhi, 2);
    // sum w: lanes with the same kq hold duplicates; reduce over kq within mr == 0
    sw += __shfl_xor_sync(0xffffffffu, sw, 1);
    sw += __shfl_xor_sync(0xffffffffu, sw, 2);
    if (kq == 0) {
        ms[wid][mr] = mlo;
        ms[wid][8 + mr] = mhi;
    }
    if (lane == 0) ms[wid][D] = sw;
    __syncthreads();
    for (int o = threadIdx.x; o < NOUT; o += MOM_THREADS) {
        double v = 0.0;
        if (o == 0) {
            for (int k = 0; k < NW; ++k) v += ms[k][D];
        } else if (o <= D) {
            for (int k = 0; k < NW; ++k) v += ms[k][o - 1];
        } else {
            // packed upper triangle -> (m, cc)
            int rem = o - 1 - D, m = 0;
            while (rem >= D - m) {
                rem -= D - m;
                ++m;
            }
            const int cc = m + rem;
            for (int k = 0; k < NW; ++k) v += cs[k][m][cc];
        }
        partials[static_cast<int64_t>(blockIdx.x) * NOUT + o] = v;
    }
}

// ---- generic d ------------------------------------------------------------------
constexpr int GEN_TILE = 64;

__global__ void __launch_bounds__(MOM_THREADS) moments_generic_kernel(const double* __restrict__ x,
                                                                      const double* __restrict__ w,
                                                                      const double* __restrict__ stats, int64_t n,
                                                                      int d, double* __restrict__ partials) {
    extern __shared__ double sm[];  // xs[GEN_TILE][d] | ws[GEN_TILE]
    double* xs = sm;
    double* ws = sm + GEN_TILE * d;
    const int nout = 1 + d + d * (d + 1) / 2;
    constexpr int MAXO = (1 + QB_MAX_D + QB_MAX_D * (QB_MAX_D + 1) / 2 + MOM_THREADS - 1) / MOM_THREADS;
    double acc[MAXO];
    int om[MAXO], oc[MAXO];  // kind of each owned output: om=-2 sum w, om=-1 mean oc, else pair (om, oc)
#pragma unroll
    for (int k = 0; k < MAXO; ++k) {
        acc[k] = 0.0;
        const int o = threadIdx.x + k * MOM_THREADS;
        om[k] = -3;
        oc[k] = 0;
        if (o < nout) {
            if (o == 0) {
                om[k] = -2;
            } else if (o <= d) {
                om[k] = -1;
                oc[k] = o - 1;
            } else {
                int rem = o - 1 - d, m = 0;
                while (rem >= d - m) {
                    rem -= d - m;
                    ++m;
                }
                om[k] = m;
                oc[k] = m + rem;
            }
        }
    }
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t ntiles = (n + GEN_TILE - 1) / GEN_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t first = t * GEN_TILE;
        const int cnt = static_cast<int>((n - first < GEN_TILE) ? (n - first) : GEN_TILE);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt * d; j += MOM_THREADS) xs[j] = ldg_stream(x + first * d + j);
        for (int j = threadIdx.x; j < cnt; j += MOM_THREADS) ws[j] = ldg_stream(w + first + j) * inv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MAXO; ++k) {
            if (om[k] == -3) continue;
            double a = acc[k];
            if (om[k] == -2) {
                for (int j = 0; j < cnt; ++j) a += ws[j];
            } else if (om[k] == -1) {
                for (int j = 0; j < cnt; ++j) a += ws[j] * xs[j * d + oc[k]];
            } else {
                for (int j = 0; j < cnt; ++j) a = fma(ws[j] * xs[j * d + om[k]], xs[j * d + oc[k]], a);
            }
            acc[k] = a;
        }
    }
#pragma unroll
    for (int k = 0; k < MAXO; ++k) {
        const int o = threadIdx.x + k * MOM_THREADS;
        if (o < nout) partials[static_cast<int64_t>(blockIdx.x) * nout + o] = acc[k];
    }
}

// ---- finish: one warp per output entry, fixed summation order, unpack the triangle ---------
__global__ void __launch_bounds__(256) moments_finish_kernel(const double* __restrict__ partials, int nblocks, int d,
                                                             double* __restrict__ out) {
    const int nout = 1 + d + d * (d + 1) / 2;
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= nout) return;
    double v = 0.0;
    for (int b = lane; b < nblocks; b += 32) v += partials[static_cast<int64_t>(b) * nout + o];
    v = warp_sum(v);
    if (lane != 0) return;
    if (o <= d) {
        out[o] = v;
    } else {
        int rem = o - 1 - d, m = 0;
        while (rem >= d - m) {
            rem -= d - m;
            ++m;
        }
        const int c = m + rem;
        out[1 + d + m * d + c] = v;
        out[1 + d + c * d + m] = v;
    }
}

static int moments_grid(int64_t n, int d) {
    const int sms = sm_count();
    int64_t want;
    if (d <= 4)
        want = (n + MOM_THREADS - 1) / MOM_THREADS;
    else if (d == 16)
        want = (n + 4 * (MOM_THREADS / 32) * 8 - 1) / (4 * (MOM_THREADS / 32) * 8);
    else
        want = (n + GEN_TILE - 1) / GEN_TILE;
    const int64_t cap = static_cast<int64_t>(sms) * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

}  // namespace qb

using namespace qb;

extern "C" size_t qb_moments_workspace_bytes(int64_t n, int32_t d) {
    (void)n;
    const size_t nout = 1 + d + static_cast<size_t>(d) * (d + 1) / 2;
    return static_cast<size_t>(256) * 4 * nout * sizeof(double) + 256;
}

extern "C" int qb_moments(const double* d_x, const double* d_w, const double* d_stats, int64_t n, int32_t d,
                          double* d_out, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_x && d_w && d_stats && d_out && d_ws, QB_ERR_INVALID_ARGUMENT, "qb_moments: NULL pointer argument");
    QB_REQUIRE(n >= 1 && d >= 1 && d <= QB_MAX_D, QB_ERR_INVALID_ARGUMENT, "qb_moments: bad n=%lld or d=%d",
               (long long)n, d);
    QB_REQUIRE(ws_bytes >= qb_moments_workspace_bytes(n, d), QB_ERR_WORKSPACE, "qb_moments: workspace too small");
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);  // [0,256) is the update ticket
    const int grid = moments_grid(n, d);
    cudaStream_t st = as_stream(stream);
    switch (d) {
        case 1: moments_small_kernel<1><<<grid, MOM_THREADS, 0, st>>>(d_x, d_w, d_stats, n, partials); break;
        case 2: moments_small_kernel<2><<<grid, MOM_THREADS, 0, st>>>(d_x, d_w, d_stats, n, partials); break;
        case 3: moments_small_kernel<3><<<grid, MOM_THREADS, 0, st>>>(d_x, d_w, d_stats, n, partials); break;
        case 4: moments_small_kernel<4><<<grid, MOM_THREADS, 0, st>>>(d_x, d_w, d_stats, n, partials); break;
        case 16: moments_d16_kernel<<<grid, MOM_THREADS, 0, st>>>(d_x, d_w, d_stats, n, partials); break;
        default: {
            const size_t smem = static_cast<size_t>(GEN_TILE) * (d + 1) * sizeof(double);
            moments_generic_kernel<<<grid, MOM_THREADS, smem, st>>>(d_x, d_w, d_stats, n, d, partials);
        }
    }
    QB_CUDA_CHECK(cudaGetLastError());
    const int nout = 1 + d + d * (d + 1) / 2;
    moments_finish_kernel<<<(nout + 7) / 8, 256, 0, st>>>(partials, grid, d, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
