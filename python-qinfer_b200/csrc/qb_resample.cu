// Liu-West resampler kernels (resamplers.py:308-372):
//   qb_cdf      — prefix sum of the normalised weights (np.cumsum)
//   qb_draw     — multinomial draw by right-bisection of the CDF (searchsorted)
//   qb_lw_move  — gather x[js], shrink towards the mean, add S @ eps, validity
//   qb_compact_invalid / qb_lw_retry — the postselection retry loop
//
// Compiled with --fmad=false (see qb_models.cuh).
#include "qb_models.cuh"

namespace qb {

// =============================================================================
// CDF
// =============================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 weights per tile

__device__ __forceinline__ double block_exclusive_scan(double v, double* warp_tot, double& block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    double base = 0.0, tot = 0.0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
        const double t = warp_tot[k];
        if (k < wid) base += t;
        tot += t;
    }
    block_total = tot;
    __syncthreads();
    return base + (inc - v);
}

// pass 1: per-tile sums of normalised weights
__global__ void __launch_bounds__(SCAN_THREADS) cdf_tile_sums_kernel(const double* __restrict__ w,
                                                                     const double* __restrict__ stats, int64_t n,
                                                                     double* __restrict__ tile_sums) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * SCAN_TILE + static_cast<int64_t>(threadIdx.x) * SCAN_ITEMS;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            if (i < n) s += w[i] * inv;
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < SCAN_THREADS / 32; ++k) tot += warp_tot[k];
            tile_sums[t] = tot;
        }
        __syncthreads();
    }
}

// pass 2: exclusive scan of the tile sums by one block (sequential chunks per thread)
__global__ void __launch_bounds__(SCAN_THREADS) cdf_scan_tiles_kernel(double* tile_sums, int64_t ntiles) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const int64_t per = (ntiles + SCAN_THREADS - 1) / SCAN_THREADS;
    const int64_t lo = static_cast<int64_t>(threadIdx.x) * per;
    const int64_t hi = (lo + per < ntiles) ? lo + per : ntiles;
    double s = 0.0;
    for (int64_t i = lo; i < hi; ++i) s += tile_sums[i];
    double total;
    double run = block_exclusive_scan(s, warp_tot, total);
    for (int64_t i = lo; i < hi; ++i) {
        const double v = tile_sums[i];
        tile_sums[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = total;  // ntiles + 1 entries: the exclusive prefix and the total
}

// pass 3: per-tile inclusive scan + tile offset
__global__ void __launch_bounds__(SCAN_THREADS) cdf_write_kernel(const double* __restrict__ w,
                                                                 const double* __restrict__ stats, int64_t n,
                                                                 const double* __restrict__ tile_offsets,
                                                                 double* __restrict__ cdf) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * SCAN_TILE + static_cast<int64_t>(threadIdx.x) * SCAN_ITEMS;
        double v[SCAN_ITEMS];
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            v[k] = (i < n) ? w[i] * inv : 0.0;
            s += v[k];
        }
        double total;
        double run = tile_offsets[t] + block_exclusive_scan(s, warp_tot, total);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            run += v[k];
            if (i < n) cdf[i] = run;
        }
    }
}

// Exact mode, first implementation: np.cumsum adds strictly left to right in
// fp64, so one lane replays that chain out of shared memory while a second warp
// streams weights in and finished CDF values out (triple-buffered chunks).
constexpr int SEQ_CHUNK = 2048;
__global__ void __launch_bounds__(64) cdf_sequential_kernel(const double* __restrict__ w,
                                                            const double* __restrict__ stats, int64_t n,
                                                            double* __restrict__ cdf) {
    __shared__ double buf[3][SEQ_CHUNK];
    const double inv = stats[QB_STAT_INV_NORM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double run = 0.0;
    const int64_t nchunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;
    for (int64_t c = 0; c < nchunks + 2; ++c) {
        if (wid == 1) {
            if (c < nchunks) {  // stage chunk c
                const int64_t first = c * SEQ_CHUNK;
                double* b = buf[c % 3];
                for (int j = lane; j < SEQ_CHUNK; j += 32) b[j] = (first + j < n) ? w[first + j] * inv : 0.0;
            }
            if (c >= 2) {  // write back chunk c-2
                const int64_t first = (c - 2) * SEQ_CHUNK;
                const double* b = buf[(c - 2) % 3];
                for (int j = lane; j < SEQ_CHUNK && first + j < n; j += 32) cdf[first + j] = b[j];
            }
        } else if (lane == 0 && c >= 1 && c <= nchunks) {  // the sequential fp64 chain on chunk c-1
            double* b = buf[(c - 1) % 3];
            const int64_t first = (c - 1) * SEQ_CHUNK;
            const int cnt = static_cast<int>((n - first < SEQ_CHUNK) ? (n - first) : SEQ_CHUNK);
            int j = 0;
            if (c == 1) {  // cumsum's first element is w[0] itself
                run = b[0];
                j = 1;
            }
#pragma unroll 8
            for (; j < cnt; ++j) {
                run = run + b[j];
                b[j] = run;
            }
        }
        __syncthreads();
    }
}

// =============================================================================
// Draw: js[i] = min(upper_bound(cdf, u[i]), n - 1)
// =============================================================================
// Plain right-bisection (24 dependent probes at n = 1e7, ~8 of them distinct random sectors).
__device__ __forceinline__ int64_t upper_bound_range(const double* __restrict__ cdf, int64_t lo, int64_t hi,
                                                     double ui) {
    while (lo < hi) {  // first index in [lo, hi] with cdf[idx] > ui  (side='right')
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= ui)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) draw_kernel(const double* __restrict__ cdf, int64_t n,
                                                   const double* __restrict__ u, int64_t n_draw,
                                                   int64_t* __restrict__ js, unsigned long long* overflow) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned int over = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_draw; i += stride) {
        int64_t lo = upper_bound_range(cdf, 0, n, ldg_stream(u + i));
        if (lo >= n) {
            lo = n - 1;
            ++over;
        }
        js[i] = lo;
    }
    if (over) atomicAdd(overflow, static_cast<unsigned long long>(over));
}

// Guide table (the classic indexed-search method for discrete sampling): g[b] = upper_bound(cdf, b / mult)
// for b = 0..M.  mult = M * 2^-e is a POWER OF TWO chosen so that cdf[n-1] * 2^-e lies in [1/2, 1): both
// u * mult and b / mult are exact in fp64, hence for b = floor(u * mult) the exact answer is bracketed,
// g[b] <= upper_bound(cdf, u) <= g[b+1], and the bisection inside the bracket returns bit-for-bit the
// same index as the full search.  With M ~ n/2..n the bracket holds one or two CDF entries, so a draw
// costs ~3 random sectors (one guide pair, one or two CDF sectors) instead of ~8.
__device__ __forceinline__ double guide_mult(const double* __restrict__ cdf, int64_t n, int64_t M) {
    int e;
    const double total = __ldg(cdf + n - 1);
    (void)frexp(total, &e);  // total = f * 2^e, f in [1/2, 1)
    return ldexp(static_cast<double>(M), -e);
}

// Each thread fills GUIDE_RUN consecutive entries: a full bisection for the first, then a gallop forward from
// the previous answer for the rest (the answers are monotone and on average less than two CDF entries apart),
// and one 32-byte store.  r1: 168 us -> the full bisection per entry was 8x too much work.
constexpr int GUIDE_RUN = 8;
__global__ void __launch_bounds__(256) guide_build_kernel(const double* __restrict__ cdf, int64_t n, int64_t M,
                                                          int32_t* __restrict__ guide) {
    const double mult = guide_mult(cdf, n, M);
    const int64_t nruns = (M + 1 + GUIDE_RUN - 1) / GUIDE_RUN;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < nruns; r += stride) {
        const int64_t b0 = r * GUIDE_RUN;
        int32_t out[GUIDE_RUN];
        int64_t idx = upper_bound_range(cdf, 0, n, static_cast<double>(b0) / mult);  // exact: mult is a power of two
        out[0] = static_cast<int32_t>(idx);
#pragma unroll
        for (int k = 1; k < GUIDE_RUN; ++k) {
            const double key = static_cast<double>(b0 + k) / mult;
            // gallop: find a bracket [idx, hi] that contains the answer, then bisect it
            int64_t step = 1, hi = idx;
            while (hi < n && __ldg(cdf + hi) <= key) {
                idx = hi + 1;
                hi += step;
                step <<= 1;
            }
            if (hi > n) hi = n;
            idx = upper_bound_range(cdf, idx, hi, key);
            out[k] = static_cast<int32_t>(idx);
        }
        if (b0 + GUIDE_RUN <= M + 1) {
            int4* dst = reinterpret_cast<int4*>(guide + b0);
            dst[0] = make_int4(out[0], out[1], out[2], out[3]);
            dst[1] = make_int4(out[4], out[5], out[6], out[7]);
        } else {
            for (int k = 0; k < GUIDE_RUN && b0 + k <= M; ++k) guide[b0 + k] = out[k];
        }
    }
}

__global__ void __launch_bounds__(256) draw_guided_kernel(const double* __restrict__ cdf, int64_t n,
                                                          const double* __restrict__ u, int64_t n_draw, int64_t M,
                                                          const int32_t* __restrict__ guide,
                                                          int64_t* __restrict__ js, unsigned long long* overflow) {
    const double mult = guide_mult(cdf, n, M);
    const double limit = static_cast<double>(M) / mult;  // = 2^e > cdf[n-1]
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned int over = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_draw; i += stride) {
        const double ui = ldg_stream(u + i);
        int64_t lo;
        if (ui >= 0.0 && ui < limit) {
            const int64_t b = static_cast<int64_t>(ui * mult);
            const int2 g = *reinterpret_cast<const int2*>(guide + (b & ~1LL));  // entries b&~1, (b&~1)+1
            int64_t glo, ghi;
            if (b & 1) {
                glo = g.y;
                ghi = __ldg(guide + b + 1);
            } else {
                glo = g.x;
                ghi = g.y;
            }
            lo = upper_bound_range(cdf, glo, ghi, ui);
        } else {
            lo = upper_bound_range(cdf, 0, n, ui);  // out-of-range or NaN uniforms: plain search
        }
        if (lo >= n) {
            lo = n - 1;
            ++over;
        }
        js[i] = lo;
    }
    if (over) atomicAdd(overflow, static_cast<unsigned long long>(over));
}

static int64_t guide_size(int64_t n) {  // M: the power of two in (n/2, n]
    int64_t m = 1;
    while (m * 2 <= n) m *= 2;
    return m;
}

// =============================================================================
// Liu-West move
// =============================================================================
struct LwParams {
    const double* x_old;
    const int64_t* js;
    const double* eps;    // (d, eps_ld) row-major
    double* x_new;
    uint8_t* invalid;
    unsigned long long* n_invalid;
    const int64_t* idxs;  // retry only
    int64_t n_old, n_new, eps_ld;
    int32_t d, tile, postselect, pad;
    double a;
    ModelView mv;
    const double* consts;  // device: S (d*d) then (1-a)*mean (d)
};

// One CTA handles `tile` consecutive new particles.
__global__ void __launch_bounds__(256) lw_move_kernel(const __grid_constant__ LwParams p) {
    extern __shared__ double sm[];
    const int d = p.d, T = p.tile, ld = d | 1;  // odd row pitch: conflict-free column walks
    double* loc = sm;                 // [T][ld]
    double* eps_s = loc + T * ld;     // [d][T]
    double* S = eps_s + d * T;        // [d][d]
    double* mshift = S + d * d;       // [d]  (1-a)*mean
    __shared__ unsigned int block_bad;
    const int tid = threadIdx.x;
    for (int j = tid; j < d * d + d; j += blockDim.x) S[j] = p.consts[j];
    const int64_t ntiles = (p.n_new + T - 1) / T;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t first = t * T;
        const int cnt = static_cast<int>((p.n_new - first < T) ? (p.n_new - first) : T);
        if (tid == 0) block_bad = 0;
        __syncthreads();
        // gather + shrink: mu = a * x[js] + (1-a) * mean   (resamplers.py:325); 16-byte loads when d is even
        if ((d & 1) == 0) {
            const int hd = d >> 1;
            for (int j = tid; j < cnt * hd; j += blockDim.x) {
                const int r = j / hd, c = (j - r * hd) * 2;
                const int64_t src = p.js[first + r];
                const double2 xv = __ldg(reinterpret_cast<const double2*>(p.x_old + src * d + c));
                loc[r * ld + c] = (p.a * xv.x) + mshift[c];
                loc[r * ld + c + 1] = (p.a * xv.y) + mshift[c + 1];
            }
        } else {
            for (int j = tid; j < cnt * d; j += blockDim.x) {
                const int r = j / d, c = j - r * d;
                const int64_t src = p.js[first + r];
                const double xv = __ldg(p.x_old + src * d + c);
                loc[r * ld + c] = (p.a * xv) + mshift[c];
            }
        }
        for (int j = tid; j < d * cnt; j += blockDim.x) {
            const int m = j / cnt, r = j - m * cnt;
            eps_s[m * T + r] = ldg_stream(p.eps + static_cast<int64_t>(m) * p.eps_ld + first + r);
        }
        __syncthreads();
        // perturb: x' = mu + (S @ eps)[:, i]   (resamplers.py:332); dgemm-style k-ordered FMA chain
        for (int j = tid; j < d * cnt; j += blockDim.x) {
            const int c = j / cnt, r = j - c * cnt;
            double z = 0.0;
            for (int m = 0; m < d; ++m) z = fma(S[c * d + m], eps_s[m * T + r], z);
            loc[r * ld + c] = loc[r * ld + c] + z;
        }
        __syncthreads();
        if (p.postselect) {
            unsigned int bad = 0;
            for (int r = tid; r < cnt; r += blockDim.x) {
                const double* xr = loc + r * ld;
                auto row = [&](int c) { return xr[c]; };
                const bool ok = model_valid(p.mv, row);
                p.invalid[first + r] = ok ? 0 : 1;
                bad += ok ? 0u : 1u;
            }
            if (bad) atomicAdd(&block_bad, bad);
        }
        for (int j = tid; j < cnt * d; j += blockDim.x) {
            const int r = j / d, c = j - r * d;
            stg_stream(p.x_new + first * d + j, loc[r * ld + c]);
        }
        __syncthreads();
        if (tid == 0 && block_bad) atomicAdd(p.n_invalid, static_cast<unsigned long long>(block_bad));
    }
}

// d <= 4: one thread per new particle, no shared-memory staging (the row fits a sector; S and the shifted mean
// travel as launch parameters, so no constant upload either).
struct LwSmallParams {
    const double* x_old;
    const int64_t* js;
    const double* eps;
    double* x_new;
    uint8_t* invalid;
    unsigned long long* n_invalid;
    int64_t n_new, eps_ld;
    int32_t postselect, pad;
    double a;
    double S[16];
    double ms[4];
    ModelView mv;
};

template <int D>
__global__ void __launch_bounds__(256) lw_move_small_kernel(const __grid_constant__ LwSmallParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int64_t nround = ((p.n_new + 31) / 32) * 32;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const bool live = i < p.n_new;
        bool ok = true;
        if (live) {
            const int64_t src = p.js[i];
            double xv[D], ev[D], out[D];
#pragma unroll
            for (int c = 0; c < D; ++c) xv[c] = __ldg(p.x_old + src * D + c);
#pragma unroll
            for (int m = 0; m < D; ++m) ev[m] = ldg_stream(p.eps + static_cast<int64_t>(m) * p.eps_ld + i);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double z = 0.0;
#pragma unroll
                for (int m = 0; m < D; ++m) z = fma(p.S[c * D + m], ev[m], z);
                out[c] = ((p.a * xv[c]) + p.ms[c]) + z;  // resamplers.py:325,332, one rounding per ufunc
            }
#pragma unroll
            for (int c = 0; c < D; ++c) stg_stream(p.x_new + i * D + c, out[c]);
            if (p.postselect) {
                auto row = [&](int c) { return out[c]; };
                ok = model_valid(p.mv, row);
                p.invalid[i] = ok ? 0 : 1;
            }
        }
        const unsigned int bad = __ballot_sync(0xffffffffu, live && !ok);
        if (bad && lane == 0) atomicAdd(p.n_invalid, static_cast<unsigned long long>(__popc(bad)));
    }
}

// Retry: one thread per still-invalid particle r (few of them).
__global__ void __launch_bounds__(128) lw_retry_kernel(const __grid_constant__ LwParams p, int64_t k) {
    const int d = p.d;
    const double* S = p.consts;
    const double* mshift = p.consts + d * d;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < k; r += stride) {
        const int64_t dst = p.idxs[r];
        // pad == 0: prefix-of-the-original-means quirk (resamplers.py:372); pad == 1: the particle's own mean
        const int64_t src = p.pad ? p.js[dst] : p.js[r];
        double xr[QB_MAX_D];
        for (int c = 0; c < d; ++c) {
            double z = 0.0;
            for (int m = 0; m < d; ++m) z = fma(S[c * d + m], p.eps[static_cast<int64_t>(m) * p.eps_ld + r], z);
            const double mu = (p.a * p.x_old[src * d + c]) + mshift[c];
            xr[c] = mu + z;
        }
        for (int c = 0; c < d; ++c) p.x_new[dst * d + c] = xr[c];
        bool ok = true;
        if (p.postselect) {
            auto row = [&](int c) { return xr[c]; };
            ok = model_valid(p.mv, row);
        }
        p.invalid[dst] = ok ? 0 : 1;
        if (!ok) atomicAdd(p.n_invalid, 1ull);
    }
}

// =============================================================================
// Ordered compaction of the invalid flags
// =============================================================================
constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = 16;
constexpr int CMP_TILE = CMP_THREADS * CMP_ITEMS;

__global__ void __launch_bounds__(CMP_THREADS) compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                    unsigned long long* __restrict__ tile_counts) {
    __shared__ unsigned int wsum[CMP_THREADS / 32];
    const int64_t ntiles = (n + CMP_TILE - 1) / CMP_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * CMP_TILE + static_cast<int64_t>(threadIdx.x) * CMP_ITEMS;
        unsigned int c = 0;
#pragma unroll
        for (int k = 0; k < CMP_ITEMS; ++k)
            if (base + k < n && flags[base + k]) ++c;
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
            for (int k = 0; k < CMP_THREADS / 32; ++k) tot += wsum[k];
            tile_counts[t] = tot;
        }
        __syncthreads();
    }
}

__global__ void compact_scan_kernel(unsigned long long* tile_counts, int64_t ntiles, int64_t* total) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long run = 0;
    for (int64_t t = 0; t < ntiles; ++t) {
        const unsigned long long c = tile_counts[t];
        tile_counts[t] = run;
        run += c;
    }
    *total = static_cast<int64_t>(run);
}

__global__ void __launch_bounds__(CMP_THREADS) compact_write_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                    const unsigned long long* __restrict__ tile_off,
                                                                    int64_t* __restrict__ out) {
    __shared__ unsigned int wsum[CMP_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t ntiles = (n + CMP_TILE - 1) / CMP_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * CMP_TILE + static_cast<int64_t>(threadIdx.x) * CMP_ITEMS;
        unsigned int c = 0;
#pragma unroll
        for (int k = 0; k < CMP_ITEMS; ++k)
            if (base + k < n && flags[base + k]) ++c;
        unsigned int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        unsigned int wbase = 0;
        for (int k = 0; k < wid; ++k) wbase += wsum[k];
        unsigned long long pos = tile_off[t] + wbase + (inc - c);
#pragma unroll
        for (int k = 0; k < CMP_ITEMS; ++k)
            if (base + k < n && flags[base + k]) out[pos++] = base + k;
        __syncthreads();
    }
}

static int capped_grid(int64_t want, int per_sm) {
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

int validate_model(const qb_model* m);

// S (d*d) and (1-a)*mean (d) live in a small device buffer owned by the library
// (one per stream would be needed for concurrent resamples; the reference path is
// single-threaded, SURVEY §8b "Threading").
static double* g_consts = nullptr;
static int g_consts_dev = -1;

static int upload_consts(const double* h_mean, const double* h_S, double a, int d, cudaStream_t st,
                         const double** out) {
    int dev = 0;
    QB_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_consts == nullptr || g_consts_dev != dev) {
        QB_CUDA_CHECK(cudaMalloc(&g_consts, (QB_MAX_D * QB_MAX_D + QB_MAX_D) * sizeof(double)));
        g_consts_dev = dev;
    }
    static thread_local double host[QB_MAX_D * QB_MAX_D + QB_MAX_D];
    for (int j = 0; j < d * d; ++j) host[j] = h_S[j];
    const double oma = 1.0 - a;
    for (int c = 0; c < d; ++c) host[d * d + c] = oma * h_mean[c];  // (1 - a) * mean
    QB_CUDA_CHECK(cudaMemcpyAsync(g_consts, host, (d * d + d) * sizeof(double), cudaMemcpyHostToDevice, st));
    *out = g_consts;
    return QB_OK;
}

}  // namespace qb

namespace qb {
size_t exact_scan_workspace_bytes(int64_t n);  // qb_scan_exact.cu
int launch_exact_scan(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, const double* tile_prefix,
                      void* d_ws, cudaStream_t st);
constexpr int64_t EXACT_PARALLEL_MIN = 1 << 15;  // below this the one-lane sequential replay is fast enough

static size_t cdf_tiles_bytes(int64_t n) {
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return ((static_cast<size_t>(ntiles + 2) * sizeof(double) + 255) / 256) * 256;
}
}  // namespace qb

using namespace qb;

extern "C" size_t qb_cdf_workspace_bytes(int64_t n) {
    return 256 + cdf_tiles_bytes(n) + exact_scan_workspace_bytes(n);
}

extern "C" int qb_cdf_exact_fallback_flag(const void* d_ws, int64_t n, int32_t* h_flag, void* stream) {
    QB_REQUIRE(d_ws && h_flag && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_cdf_exact_fallback_flag: bad arguments");
    const unsigned char* base = reinterpret_cast<const unsigned char*>(d_ws) + 256 + cdf_tiles_bytes(n);
    QB_CUDA_CHECK(cudaMemcpyAsync(h_flag, base + 64, sizeof(int32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
    QB_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    return QB_OK;
}

extern "C" int qb_cdf(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, int32_t mode, void* d_ws,
                      size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_w && d_stats && d_cdf && d_ws && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_cdf: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_cdf_workspace_bytes(n), QB_ERR_WORKSPACE, "qb_cdf: workspace too small");
    cudaStream_t st = as_stream(stream);
    if (mode == QB_SCAN_EXACT && n < EXACT_PARALLEL_MIN) {
        cdf_sequential_kernel<<<1, 64, 0, st>>>(d_w, d_stats, n, d_cdf);
        QB_CUDA_CHECK(cudaGetLastError());
        return QB_OK;
    }
    QB_REQUIRE(mode == QB_SCAN_FAST || mode == QB_SCAN_EXACT, QB_ERR_INVALID_ARGUMENT, "qb_cdf: unknown mode %d", mode);
    double* tiles = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);  // [0,256) is the update ticket
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int grid = capped_grid(ntiles, 8);
    cdf_tile_sums_kernel<<<grid, SCAN_THREADS, 0, st>>>(d_w, d_stats, n, tiles);
    QB_CUDA_CHECK(cudaGetLastError());
    cdf_scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(tiles, ntiles);
    QB_CUDA_CHECK(cudaGetLastError());
    if (mode == QB_SCAN_EXACT) {
        // the approximate tile prefix predicts the binade of every segment; the exact values come from the replay scan
        return launch_exact_scan(d_w, d_stats, n, d_cdf, tiles,
                                 reinterpret_cast<unsigned char*>(d_ws) + 256 + cdf_tiles_bytes(n), st);
    }
    cdf_write_kernel<<<grid, SCAN_THREADS, 0, st>>>(d_w, d_stats, n, tiles, d_cdf);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" size_t qb_draw_workspace_bytes(int64_t n) {
    if (n < 4096 || n >= (1LL << 31)) return 0;  // small clouds / int32 overflow: plain bisection
    return static_cast<size_t>(guide_size(n) + 2) * sizeof(int32_t) + 256;
}

extern "C" int qb_draw(const double* d_cdf, int64_t n, const double* d_u, int64_t n_draw, int64_t* d_js,
                       int64_t* d_overflow, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_cdf && d_u && d_js && d_overflow && n >= 1 && n_draw >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_draw: bad arguments");
    cudaStream_t st = as_stream(stream);
    QB_CUDA_CHECK(cudaMemsetAsync(d_overflow, 0, sizeof(int64_t), st));
    const size_t need = qb_draw_workspace_bytes(n);
    const int grid = capped_grid((n_draw + 255) / 256, 8);
    // the guide pays for itself when there are at least about as many draws as table entries
    if (d_ws != nullptr && need > 0 && ws_bytes >= need && n_draw * 4 >= n) {
        const int64_t M = guide_size(n);
        int32_t* guide = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
        guide_build_kernel<<<capped_grid((M / GUIDE_RUN + 256) / 256, 8), 256, 0, st>>>(d_cdf, n, M, guide);
        QB_CUDA_CHECK(cudaGetLastError());
        draw_guided_kernel<<<grid, 256, 0, st>>>(d_cdf, n, d_u, n_draw, M, guide, d_js,
                                                  reinterpret_cast<unsigned long long*>(d_overflow));
    } else {
        draw_kernel<<<grid, 256, 0, st>>>(d_cdf, n, d_u, n_draw, d_js,
                                          reinterpret_cast<unsigned long long*>(d_overflow));
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d, const int64_t* d_js,
                          const double* h_mean, const double* h_S, double a, const double* d_eps, int64_t n_new,
                          double* d_x_new, int32_t postselect, uint8_t* d_invalid, int64_t* d_n_invalid,
                          void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x_old && d_js && h_mean && h_S && d_eps && d_x_new && d_invalid && d_n_invalid,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_move: NULL pointer argument");
    QB_REQUIRE(d == model->d && n_old >= 1 && n_new >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_move: bad sizes");
    cudaStream_t st = as_stream(stream);
    if (d <= 4) {
        LwSmallParams q;
        q.x_old = d_x_old;
        q.js = d_js;
        q.eps = d_eps;
        q.x_new = d_x_new;
        q.invalid = d_invalid;
        q.n_invalid = reinterpret_cast<unsigned long long*>(d_n_invalid);
        q.n_new = n_new;
        q.eps_ld = n_new;
        q.postselect = postselect;
        q.pad = 0;
        q.a = a;
        for (int j = 0; j < 16; ++j) q.S[j] = (j < d * d) ? h_S[j] : 0.0;
        const double oma = 1.0 - a;
        for (int c = 0; c < 4; ++c) q.ms[c] = (c < d) ? oma * h_mean[c] : 0.0;  // (1 - a) * mean
        q.mv = make_model_view(*model);
        QB_CUDA_CHECK(cudaMemsetAsync(d_n_invalid, 0, sizeof(int64_t), st));
        if (!postselect) QB_CUDA_CHECK(cudaMemsetAsync(d_invalid, 0, static_cast<size_t>(n_new), st));
        const int grid = capped_grid((n_new + 255) / 256, 8);
        switch (d) {
            case 1: lw_move_small_kernel<1><<<grid, 256, 0, st>>>(q); break;
            case 2: lw_move_small_kernel<2><<<grid, 256, 0, st>>>(q); break;
            case 3: lw_move_small_kernel<3><<<grid, 256, 0, st>>>(q); break;
            default: lw_move_small_kernel<4><<<grid, 256, 0, st>>>(q); break;
        }
        QB_CUDA_CHECK(cudaGetLastError());
        return QB_OK;
    }
    LwParams p;
    rc = upload_consts(h_mean, h_S, a, d, st, &p.consts);
    if (rc != QB_OK) return rc;
    p.x_old = d_x_old;
    p.js = d_js;
    p.eps = d_eps;
    p.x_new = d_x_new;
    p.invalid = d_invalid;
    p.n_invalid = reinterpret_cast<unsigned long long*>(d_n_invalid);
    p.idxs = nullptr;
    p.n_old = n_old;
    p.n_new = n_new;
    p.eps_ld = n_new;
    p.d = d;
    p.tile = (d <= 4) ? 512 : ((d <= 16) ? 128 : 32);
    p.postselect = postselect;
    p.pad = 0;
    p.a = a;
    p.mv = make_model_view(*model);
    QB_CUDA_CHECK(cudaMemsetAsync(d_n_invalid, 0, sizeof(int64_t), st));
    if (!postselect) QB_CUDA_CHECK(cudaMemsetAsync(d_invalid, 0, static_cast<size_t>(n_new), st));
    const size_t smem = (static_cast<size_t>(p.tile) * (d | 1) + static_cast<size_t>(d) * p.tile + d * d + d) *
                        sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        QB_CUDA_CHECK(cudaFuncSetAttribute(lw_move_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    const int64_t ntiles = (n_new + p.tile - 1) / p.tile;
    lw_move_kernel<<<capped_grid(ntiles, 6), 256, smem, st>>>(p);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" size_t qb_compact_workspace_bytes(int64_t n) {
    const int64_t ntiles = (n + CMP_TILE - 1) / CMP_TILE;
    return static_cast<size_t>(ntiles + 1) * sizeof(unsigned long long) + 256;
}

extern "C" int qb_compact_invalid(const uint8_t* d_invalid, int64_t n, int64_t* d_idxs_out, int64_t* d_count,
                                  void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_invalid && d_idxs_out && d_count && d_ws && n >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_compact_invalid: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_compact_workspace_bytes(n), QB_ERR_WORKSPACE, "qb_compact_invalid: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned long long* tiles = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    const int64_t ntiles = (n + CMP_TILE - 1) / CMP_TILE;
    const int grid = capped_grid(ntiles, 8);
    compact_count_kernel<<<grid, CMP_THREADS, 0, st>>>(d_invalid, n, tiles);
    QB_CUDA_CHECK(cudaGetLastError());
    compact_scan_kernel<<<1, 32, 0, st>>>(tiles, ntiles, d_count);
    QB_CUDA_CHECK(cudaGetLastError());
    compact_write_kernel<<<grid, CMP_THREADS, 0, st>>>(d_invalid, n, tiles, d_idxs_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d, const int64_t* d_js,
                           const int64_t* d_idxs, int64_t k, const double* h_mean, const double* h_S, double a,
                           const double* d_eps, double* d_x_new, uint8_t* d_invalid, int64_t* d_n_invalid,
                           int32_t own_mean, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x_old && d_js && d_idxs && h_mean && h_S && d_eps && d_x_new && d_invalid && d_n_invalid,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_retry: NULL pointer argument");
    QB_REQUIRE(d == model->d && n_old >= 1 && k >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_retry: bad sizes");
    cudaStream_t st = as_stream(stream);
    LwParams p;
    rc = upload_consts(h_mean, h_S, a, d, st, &p.consts);
    if (rc != QB_OK) return rc;
    p.x_old = d_x_old;
    p.js = d_js;
    p.eps = d_eps;
    p.x_new = d_x_new;
    p.invalid = d_invalid;
    p.n_invalid = reinterpret_cast<unsigned long long*>(d_n_invalid);
    p.idxs = d_idxs;
    p.n_old = n_old;
    p.n_new = 0;
    p.eps_ld = k;
    p.d = d;
    p.tile = 0;
    p.postselect = 1;
    p.pad = own_mean ? 1 : 0;
    p.a = a;
    p.mv = make_model_view(*model);
    QB_CUDA_CHECK(cudaMemsetAsync(d_n_invalid, 0, sizeof(int64_t), st));
    lw_retry_kernel<<<capped_grid((k + 127) / 128, 8), 128, 0, st>>>(p, k);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
