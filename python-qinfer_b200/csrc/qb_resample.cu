// Liu-West resampler kernels (resamplers.py:308-372):
//   qb_cdf      — prefix sum of the normalised weights (np.cumsum)
//   qb_draw     — multinomial draw by right-bisection of the CDF (searchsorted)
//   qb_lw_move  — gather x[js], shrink towards the mean, add S @ eps, validity
//   qb_compact_invalid / qb_lw_retry — the postselection retry loop
//
// Compiled with --fmad=false (see qb_models.cuh).
#include "qb_models.cuh"
#include "qb_philox.cuh"

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

// Guide table header, stored in front of a table built by cdf_write_kernel<true> (the fused draw+move kernels read it).
struct GuideHdr {
    double mult;   // bucket of u = floor(u * mult); a power of two, so u * mult and b / mult are exact
    double limit;  // = M / mult: uniforms >= limit (or < 0, NaN) take the plain bisection
    int64_t M;     // table entries 0..M; 0: no table (small clouds), plain bisection everywhere
    int64_t pad;
};

// mult for a table of M buckets covering uniforms in [0, range): the smallest power of two 2^e >= range.
__device__ __forceinline__ void guide_scale(double range, int64_t M, double& mult, double& limit) {
    int e;
    const double f = frexp(range, &e);  // range = f * 2^e, f in [1/2, 1)
    if (f == 0.5) --e;                  // range is itself a power of two
    mult = ldexp(static_cast<double>(M), -e);
    limit = ldexp(1.0, e);
}

// bucket boundary of a stored CDF value: B(c) = ceil(c * mult) clamped to [0, M + 1].  Entry i owns the buckets
// [B(cdf[i-1]), B(cdf[i])): exactly the b with cdf[i-1] <= b / mult < cdf[i], i.e. upper_bound(cdf, b / mult) = i.
__device__ __forceinline__ int64_t guide_boundary(double c, double mult, int64_t M) {
    const double y = ceil(c * mult);
    if (!(y > 0.0)) return 0;  // negative or NaN
    if (y > static_cast<double>(M)) return M + 1;
    return static_cast<int64_t>(y);
}

constexpr int GUIDE_Q = 64;           // per-block queue of long bucket ranges (one heavy particle), filled cooperatively
constexpr int GUIDE_SHORT_RANGE = 16;

// pass 3: per-tile inclusive scan + tile offset.  The last entry of every full tile is stored as the next tile's
// offset, so that the value a tile starts from IS the stored predecessor (needed by the guide scatter; the two
// differ only in the rounding of the last bit).
// GUIDE: while the CDF values sit in registers, scatter the guide table g[b] = upper_bound(cdf, b / mult) that the
// draw kernels use to bracket the bisection (replaces a separate pass of M bisections over the finished CDF).
template <bool GUIDE>
__global__ void __launch_bounds__(SCAN_THREADS) cdf_write_kernel(const double* __restrict__ w,
                                                                 const double* __restrict__ stats, int64_t n,
                                                                 const double* __restrict__ tile_offsets,
                                                                 double* __restrict__ cdf, int64_t M, int scaled,
                                                                 GuideHdr* __restrict__ hdr,
                                                                 int32_t* __restrict__ guide) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    __shared__ long long warp_lastB[SCAN_THREADS / 32];
    __shared__ long long q_lo[GUIDE_Q], q_hi[GUIDE_Q];
    __shared__ int q_val[GUIDE_Q];
    __shared__ int q_n;
    const double inv = stats[QB_STAT_INV_NORM];
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double mult = 0.0, limit = 0.0;
    int64_t b_top = 0;
    if (GUIDE) {
        // range of the uniforms the table serves: [0, 1), or [0, total) for a shard that scales its draws
        const double total_hi = tile_offsets[ntiles] * (1.0 + 9.5367431640625e-07);  // (1 + 2^-20): above cdf[n-1]
        guide_scale(scaled ? total_hi : 1.0, M, mult, limit);
        b_top = guide_boundary(total_hi, mult, M);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            hdr->mult = mult;
            hdr->limit = limit;
            hdr->M = M;
            hdr->pad = 0;
        }
        if (threadIdx.x == 0) q_n = 0;
        // buckets at or above the total hold n
        const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
        for (int64_t b = b_top + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; b <= M; b += stride)
            guide[b] = static_cast<int32_t>(n);
        __syncthreads();
    }
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
        const double off = tile_offsets[t];
        double run = off + block_exclusive_scan(s, warp_tot, total);
        double c[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            run += v[k];
            c[k] = run;
        }
        if (threadIdx.x == SCAN_THREADS - 1 && t + 1 < ntiles) c[SCAN_ITEMS - 1] = tile_offsets[t + 1];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) cdf[base + k] = c[k];
        if (GUIDE) {
            int64_t B[SCAN_ITEMS];
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) B[k] = guide_boundary(c[k], mult, M);
            // boundary of the stored predecessor of my first entry: previous thread's last entry (shuffle / shared),
            // or the tile offset itself for the first thread (= the previous tile's stored last entry; 0 for tile 0)
            long long prevB = __shfl_up_sync(0xffffffffu, static_cast<long long>(B[SCAN_ITEMS - 1]), 1);
            if (lane == 31) warp_lastB[wid] = B[SCAN_ITEMS - 1];
            __syncthreads();
            if (lane == 0) prevB = (wid == 0) ? guide_boundary(off, mult, M) : warp_lastB[wid - 1];
            int64_t lo = prevB;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                const int64_t i = base + k;
                if (i < n) {
                    int64_t hi = B[k];
                    const int32_t val = static_cast<int32_t>(i);
                    if (hi - lo > GUIDE_SHORT_RANGE) {
                        const int slot = atomicAdd(&q_n, 1);
                        if (slot < GUIDE_Q) {
                            q_lo[slot] = lo;
                            q_hi[slot] = hi;
                            q_val[slot] = val;
                            hi = lo;  // handed over to the block
                        }
                    }
                    for (int64_t b = lo; b < hi; ++b) guide[b] = val;
                    lo = (B[k] > lo) ? B[k] : lo;
                    if (i == n - 1)  // between the actual total and the conservative b_top: still "past the end"
                        for (int64_t b = lo; b < b_top; ++b) guide[b] = static_cast<int32_t>(n);
                }
            }
            __syncthreads();
            const int nq = (q_n < GUIDE_Q) ? q_n : GUIDE_Q;
            for (int e = 0; e < nq; ++e) {
                const int32_t val = q_val[e];
                for (int64_t b = q_lo[e] + threadIdx.x; b < q_hi[e]; b += SCAN_THREADS) guide[b] = val;
            }
            __syncthreads();
            if (threadIdx.x == 0) q_n = 0;
            __syncthreads();
        }
    }
}

// Exact mode, first implementation: np.cumsum adds strictly left to right in
// fp64, so one lane replays that chain out of shared memory while a second warp
// streams weights in and finished CDF values out (triple-buffered chunks).
constexpr int SEQ_CHUNK = 2048;
__global__ void __launch_bounds__(64) cdf_sequential_kernel(const double* __restrict__ w,
                                                            const double* __restrict__ stats, int64_t n,
                                                            double* __restrict__ cdf,
                                                            const double* __restrict__ carry) {
    __shared__ double buf[3][SEQ_CHUNK];
    const double inv = stats[QB_STAT_INV_NORM];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double run = carry ? *carry : 0.0;  // (chained scan: the running rounded sum of the slabs before this one)
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
            if (c == 1 && !carry) {  // cumsum's first element is w[0] itself
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

// -----------------------------------------------------------------------------
// Fused draw + move for the device-RNG mode (d <= 4): one pass does what rng_kernel<uniform>, draw_guided_kernel,
// rng_kernel<normal> and lw_move_small_kernel do in four, without materialising u, js or eps: each thread
// regenerates the Philox variates of two consecutive new particles (bit-identical to the stand-alone generators),
// brackets the bisection with the guide table cdf_write_kernel<true> scattered, gathers the parent row, shrinks,
// perturbs, tests validity and writes the new row.  The retry kernel regenerates the parent index the same way.
// -----------------------------------------------------------------------------
struct LwFusedParams {
    const double* x_old;
    const double* cdf;
    const GuideHdr* hdr;     // NULL: plain bisection
    const int32_t* guide;
    double* x_new;
    uint8_t* invalid;
    unsigned long long* counters;  // [0] invalid, [1] clamped draws
    const int64_t* idxs;           // retry only
    int64_t n_old, n_new, k;
    uint64_t seed_u, off_u, seed_n, off_n;
    int32_t postselect, scale_u, own_mean, pad;
    double a;
    double S[16];
    double ms[4];
    ModelView mv;
};

struct DrawCtx {
    const double* cdf;
    const int32_t* guide;
    int64_t n, M;
    double mult, limit, scale;
};

__device__ __forceinline__ DrawCtx make_draw_ctx(const LwFusedParams& p) {
    DrawCtx c;
    c.cdf = p.cdf;
    c.guide = p.guide;
    c.n = p.n_old;
    c.M = 0;
    c.mult = 0.0;
    c.limit = 0.0;
    if (p.hdr != nullptr) {
        c.M = p.hdr->M;
        c.mult = p.hdr->mult;
        c.limit = p.hdr->limit;
    }
    c.scale = p.scale_u ? __ldg(p.cdf + p.n_old - 1) : 1.0;
    return c;
}

// min(upper_bound(cdf, u), n - 1); `over` counts the clamps
__device__ __forceinline__ int64_t guided_draw(const DrawCtx& c, double u, unsigned int& over) {
    const double ui = c.scale == 1.0 ? u : u * c.scale;
    int64_t lo;
    if (c.M > 0 && ui >= 0.0 && ui < c.limit) {
        const int64_t b = static_cast<int64_t>(ui * c.mult);
        const int2 g = *reinterpret_cast<const int2*>(c.guide + (b & ~1LL));
        int64_t glo, ghi;
        if (b & 1) {
            glo = g.y;
            ghi = __ldg(c.guide + b + 1);
        } else {
            glo = g.x;
            ghi = g.y;
        }
        lo = upper_bound_range(c.cdf, glo, ghi, ui);
    } else {
        lo = upper_bound_range(c.cdf, 0, c.n, ui);
    }
    if (lo >= c.n) {
        lo = c.n - 1;
        ++over;
    }
    return lo;
}

template <int D>
__global__ void __launch_bounds__(256) lw_draw_move_kernel(const __grid_constant__ LwFusedParams p) {
    const DrawCtx dc = make_draw_ctx(p);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int64_t npairs = (p.n_new + 1) / 2;
    const int64_t nround = ((npairs + 31) / 32) * 32;
    unsigned int over = 0;
    for (int64_t pr = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pr < nround; pr += stride) {
        const int64_t i0 = 2 * pr;
        unsigned int nbad = 0;
        if (i0 < p.n_new) {
            const bool two = i0 + 1 < p.n_new;
            double u[2];
            philox_uniform_pair(p.seed_u, p.off_u + static_cast<uint64_t>(pr), u[0], u[1]);
            int64_t src[2];
            src[0] = guided_draw(dc, u[0], over);
            src[1] = two ? guided_draw(dc, u[1], over) : src[0];
            double xv[2][D];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < D; ++c) xv[s][c] = __ldg(p.x_old + src[s] * D + c);
            // eps[m][i] is element m * n_new + i of the normal stream (the (d, k) row-major layout of kernel(n_rvs, k))
            double ev[2][D];
#pragma unroll
            for (int m = 0; m < D; ++m) {
                const int64_t f0 = static_cast<int64_t>(m) * p.n_new + i0;
                if ((f0 & 1) == 0) {
                    philox_normal_pair(p.seed_n, p.off_n + static_cast<uint64_t>(f0 >> 1), ev[0][m], ev[1][m]);
                } else {
                    ev[0][m] = philox_normal_elem(p.seed_n, p.off_n, f0);
                    ev[1][m] = two ? philox_normal_elem(p.seed_n, p.off_n, f0 + 1) : 0.0;
                }
            }
            double out[2][D];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double z = 0.0;
#pragma unroll
                    for (int m = 0; m < D; ++m) z = fma(p.S[c * D + m], ev[s][m], z);
                    out[s][c] = ((p.a * xv[s][c]) + p.ms[c]) + z;  // resamplers.py:325,332, one rounding per ufunc
                }
            double* dst = p.x_new + i0 * D;
            if (D == 1 && two && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(out[0][0]),
                             "d"(out[1][0])
                             : "memory");
            } else {
#pragma unroll
                for (int c = 0; c < D; ++c) stg_stream(dst + c, out[0][c]);
                if (two) {
#pragma unroll
                    for (int c = 0; c < D; ++c) stg_stream(dst + D + c, out[1][c]);
                }
            }
            if (p.postselect) {
                auto row0 = [&](int c) { return out[0][c]; };
                auto row1 = [&](int c) { return out[1][c]; };
                const bool ok0 = model_valid(p.mv, row0);
                const bool ok1 = two ? model_valid(p.mv, row1) : true;
                p.invalid[i0] = ok0 ? 0 : 1;
                if (two) p.invalid[i0 + 1] = ok1 ? 0 : 1;
                nbad = (ok0 ? 0u : 1u) + (ok1 ? 0u : 1u);
            }
        }
        const unsigned int tot = __reduce_add_sync(0xffffffffu, nbad);
        if (tot && lane == 0) atomicAdd(p.counters, static_cast<unsigned long long>(tot));
    }
    if (over) atomicAdd(p.counters + 1, static_cast<unsigned long long>(over));
}

// Retry of the k still-invalid particles idxs[0..k): parent = draw(u[r]) (the reference's prefix-of-the-original-means
// quirk, resamplers.py:372) or draw(u[idxs[r]]) (own_mean), fresh normals eps[m][r] = element m * k + r of this
// iteration's normal stream.
template <int D>
__global__ void __launch_bounds__(128) lw_draw_retry_kernel(const __grid_constant__ LwFusedParams p) {
    const DrawCtx dc = make_draw_ctx(p);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned int over = 0;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.k; r += stride) {
        const int64_t dst = p.idxs[r];
        const int64_t slot = p.own_mean ? dst : r;
        const int64_t src = guided_draw(dc, philox_uniform_elem(p.seed_u, p.off_u, slot), over);
        double out[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double z = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m)
                z = fma(p.S[c * D + m], philox_normal_elem(p.seed_n, p.off_n, static_cast<int64_t>(m) * p.k + r), z);
            out[c] = ((p.a * __ldg(p.x_old + src * D + c)) + p.ms[c]) + z;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) p.x_new[dst * D + c] = out[c];
        auto row = [&](int c) { return out[c]; };
        const bool ok = model_valid(p.mv, row);
        p.invalid[dst] = ok ? 0 : 1;
        if (!ok) atomicAdd(p.counters, 1ull);
    }
    (void)over;  // clamps were counted by the first pass
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

// exclusive scan of the tile counts by one block: sequential chunks per thread + a block scan of the chunk sums
// (the counts are 64-bit integers: any association order gives the same result)
__global__ void __launch_bounds__(CMP_THREADS) compact_scan_kernel(unsigned long long* tile_counts, int64_t ntiles,
                                                                   int64_t* total) {
    __shared__ unsigned long long wsum[CMP_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t per = (ntiles + CMP_THREADS - 1) / CMP_THREADS;
    const int64_t lo = static_cast<int64_t>(threadIdx.x) * per;
    const int64_t hi = (lo + per < ntiles) ? lo + per : ntiles;
    unsigned long long s = 0;
    for (int64_t t = lo; t < hi; ++t) s += tile_counts[t];
    unsigned long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < CMP_THREADS / 32; ++k) {
        if (k < wid) base += wsum[k];
        tot += wsum[k];
    }
    unsigned long long run = base + (inc - s);
    for (int64_t t = lo; t < hi; ++t) {
        const unsigned long long c = tile_counts[t];
        tile_counts[t] = run;
        run += c;
    }
    if (threadIdx.x == 0) *total = static_cast<int64_t>(tot);
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

// S (d*d) and (1-a)*mean (d) of the generic-d kernels live in a CALLER-owned device buffer of
// qb_lw_move_workspace_bytes(d) bytes (one per cloud: two updaters on two streams do not share anything)
static int upload_consts(const double* h_mean, const double* h_S, double a, int d, cudaStream_t st, void* d_ws,
                         size_t ws_bytes, const double** out) {
    QB_REQUIRE(d_ws != nullptr && ws_bytes >= static_cast<size_t>(d) * (d + 1) * sizeof(double), QB_ERR_WORKSPACE,
               "qb_lw_move/retry: n_modelparams > 4 needs a workspace of qb_lw_move_workspace_bytes(d) bytes");
    static thread_local double host[QB_MAX_D * QB_MAX_D + QB_MAX_D];
    for (int j = 0; j < d * d; ++j) host[j] = h_S[j];
    const double oma = 1.0 - a;
    for (int c = 0; c < d; ++c) host[d * d + c] = oma * h_mean[c];  // (1 - a) * mean
    QB_CUDA_CHECK(cudaMemcpyAsync(d_ws, host, (d * d + d) * sizeof(double), cudaMemcpyHostToDevice, st));
    *out = reinterpret_cast<const double*>(d_ws);
    return QB_OK;
}

}  // namespace qb

namespace qb {
size_t exact_scan_workspace_bytes(int64_t n);  // qb_scan_exact.cu
int launch_exact_scan(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, const double* tile_prefix,
                      const double* d_carry, void* d_ws, cudaStream_t st);
constexpr int64_t EXACT_PARALLEL_MIN = 1 << 15;  // below this the one-lane sequential replay is fast enough

static size_t cdf_tiles_bytes(int64_t n) {
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return ((static_cast<size_t>(ntiles + 2) * sizeof(double) + 255) / 256) * 256;
}
}  // namespace qb

using namespace qb;

// Workspace layout of the CDF family: [0,256) update ticket | tile sums | exact-scan scratch | guide header + table.
static size_t guide_offset(int64_t n) {
    return ((256 + cdf_tiles_bytes(n) + exact_scan_workspace_bytes(n) + 255) / 256) * 256;
}
static bool guide_wanted(int64_t n) { return n >= 4096 && n < (1LL << 31); }

extern "C" size_t qb_cdf_workspace_bytes(int64_t n) {
    size_t b = guide_offset(n) + 64;
    if (guide_wanted(n)) b += static_cast<size_t>(guide_size(n) + 2) * sizeof(int32_t);
    return b + 256;
}

extern "C" int qb_cdf_exact_fallback_flag(const void* d_ws, int64_t n, int32_t* h_flag, void* stream) {
    QB_REQUIRE(d_ws && h_flag && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_cdf_exact_fallback_flag: bad arguments");
    const unsigned char* base = reinterpret_cast<const unsigned char*>(d_ws) + 256 + cdf_tiles_bytes(n);
    QB_CUDA_CHECK(cudaMemcpyAsync(h_flag, base + 64, sizeof(int32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
    QB_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    return QB_OK;
}

static int cdf_impl(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, int32_t mode,
                    const double* d_carry, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_w && d_stats && d_cdf && d_ws && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_cdf: bad arguments");
    QB_REQUIRE(ws_bytes >= qb_cdf_workspace_bytes(n), QB_ERR_WORKSPACE, "qb_cdf: workspace too small");
    cudaStream_t st = as_stream(stream);
    if (mode == QB_SCAN_EXACT && n < EXACT_PARALLEL_MIN) {
        cdf_sequential_kernel<<<1, 64, 0, st>>>(d_w, d_stats, n, d_cdf, d_carry);
        QB_CUDA_CHECK(cudaGetLastError());
        return QB_OK;
    }
    QB_REQUIRE(mode == QB_SCAN_FAST || mode == QB_SCAN_EXACT || mode == QB_SCAN_FAST_GUIDE ||
                   mode == QB_SCAN_FAST_GUIDE_SCALED,
               QB_ERR_INVALID_ARGUMENT, "qb_cdf: unknown mode %d", mode);
    double* tiles = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);  // [0,256) is the update ticket
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int grid = capped_grid(ntiles, 8);
    cdf_tile_sums_kernel<<<grid, SCAN_THREADS, 0, st>>>(d_w, d_stats, n, tiles);
    QB_CUDA_CHECK(cudaGetLastError());
    cdf_scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(tiles, ntiles);
    QB_CUDA_CHECK(cudaGetLastError());
    if (mode == QB_SCAN_EXACT) {
        // the approximate tile prefix predicts the binade of every segment; the exact values come from the replay scan
        return launch_exact_scan(d_w, d_stats, n, d_cdf, tiles, d_carry,
                                 reinterpret_cast<unsigned char*>(d_ws) + 256 + cdf_tiles_bytes(n), st);
    }
    GuideHdr* hdr = reinterpret_cast<GuideHdr*>(reinterpret_cast<unsigned char*>(d_ws) + guide_offset(n));
    int32_t* guide = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(hdr) + 64);
    if (mode == QB_SCAN_FAST || !guide_wanted(n)) {
        if (mode != QB_SCAN_FAST) QB_CUDA_CHECK(cudaMemsetAsync(hdr, 0, sizeof(GuideHdr), st));  // M = 0: no table
        cdf_write_kernel<false><<<grid, SCAN_THREADS, 0, st>>>(d_w, d_stats, n, tiles, d_cdf, 0, 0, nullptr, nullptr);
    } else {
        cdf_write_kernel<true><<<grid, SCAN_THREADS, 0, st>>>(d_w, d_stats, n, tiles, d_cdf, guide_size(n),
                                                              mode == QB_SCAN_FAST_GUIDE_SCALED ? 1 : 0, hdr, guide);
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_cdf(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, int32_t mode, void* d_ws,
                      size_t ws_bytes, void* stream) {
    return cdf_impl(d_w, d_stats, n, d_cdf, mode, nullptr, d_ws, ws_bytes, stream);
}

extern "C" int qb_cdf_chained(const double* d_w, const double* d_stats, int64_t n, double* d_cdf,
                              const double* d_carry_in, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_carry_in, QB_ERR_INVALID_ARGUMENT, "qb_cdf_chained: d_carry_in is NULL");
    return cdf_impl(d_w, d_stats, n, d_cdf, QB_SCAN_EXACT, d_carry_in, d_ws, ws_bytes, stream);
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
                          void* d_ws, size_t ws_bytes, void* stream) {
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
    rc = upload_consts(h_mean, h_S, a, d, st, d_ws, ws_bytes, &p.consts);
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

extern "C" size_t qb_lw_move_workspace_bytes(int32_t d) {
    return static_cast<size_t>(d) * (d + 1) * sizeof(double);
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
    compact_scan_kernel<<<1, CMP_THREADS, 0, st>>>(tiles, ntiles, d_count);
    QB_CUDA_CHECK(cudaGetLastError());
    compact_write_kernel<<<grid, CMP_THREADS, 0, st>>>(d_invalid, n, tiles, d_idxs_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d, const int64_t* d_js,
                           const int64_t* d_idxs, int64_t k, const double* h_mean, const double* h_S, double a,
                           const double* d_eps, double* d_x_new, uint8_t* d_invalid, int64_t* d_n_invalid,
                           int32_t own_mean, void* d_ws, size_t ws_bytes, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x_old && d_js && d_idxs && h_mean && h_S && d_eps && d_x_new && d_invalid && d_n_invalid,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_retry: NULL pointer argument");
    QB_REQUIRE(d == model->d && n_old >= 1 && k >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_retry: bad sizes");
    cudaStream_t st = as_stream(stream);
    LwParams p;
    rc = upload_consts(h_mean, h_S, a, d, st, d_ws, ws_bytes, &p.consts);
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

// ---- fused draw + move (device-RNG mode, d <= 4) ------------------------------------------------------------
static int fill_fused(LwFusedParams& q, const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                      const double* d_cdf, const void* d_ws, size_t ws_bytes, int32_t use_guide, const double* h_mean,
                      const double* h_S, double a, uint64_t seed_u, uint64_t off_u, uint64_t seed_n, uint64_t off_n,
                      int32_t scale_u, double* d_x_new, uint8_t* d_invalid, int64_t* d_counters) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x_old && d_cdf && h_mean && h_S && d_x_new && d_invalid && d_counters, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_draw_move/retry: NULL pointer argument");
    QB_REQUIRE(d == model->d && d >= 1 && d <= 4 && n_old >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_draw_move/retry: needs 1 <= d <= 4 (larger d: qb_draw + qb_lw_move)");
    q.x_old = d_x_old;
    q.cdf = d_cdf;
    q.hdr = nullptr;
    q.guide = nullptr;
    if (use_guide) {
        QB_REQUIRE(d_ws && ws_bytes >= qb_cdf_workspace_bytes(n_old), QB_ERR_WORKSPACE,
                   "qb_lw_draw_move/retry: workspace too small for the guide table");
        const unsigned char* base = reinterpret_cast<const unsigned char*>(d_ws) + guide_offset(n_old);
        q.hdr = reinterpret_cast<const GuideHdr*>(base);
        q.guide = reinterpret_cast<const int32_t*>(base + 64);
    }
    q.x_new = d_x_new;
    q.invalid = d_invalid;
    q.counters = reinterpret_cast<unsigned long long*>(d_counters);
    q.idxs = nullptr;
    q.n_old = n_old;
    q.n_new = 0;
    q.k = 0;
    q.seed_u = seed_u;
    q.off_u = off_u;
    q.seed_n = seed_n;
    q.off_n = off_n;
    q.postselect = 1;
    q.scale_u = scale_u ? 1 : 0;
    q.own_mean = 0;
    q.pad = 0;
    q.a = a;
    for (int j = 0; j < 16; ++j) q.S[j] = (j < d * d) ? h_S[j] : 0.0;
    const double oma = 1.0 - a;
    for (int c = 0; c < 4; ++c) q.ms[c] = (c < d) ? oma * h_mean[c] : 0.0;  // (1 - a) * mean
    q.mv = make_model_view(*model);
    return QB_OK;
}

extern "C" int qb_lw_draw_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                               const double* d_cdf, const void* d_ws, size_t ws_bytes, int32_t use_guide,
                               const double* h_mean, const double* h_S, double a, uint64_t seed_u, uint64_t off_u,
                               uint64_t seed_n, uint64_t off_n, int32_t scale_u, int64_t n_new, double* d_x_new,
                               int32_t postselect, uint8_t* d_invalid, int64_t* d_counters, void* stream) {
    LwFusedParams q;
    int rc = fill_fused(q, model, d_x_old, n_old, d, d_cdf, d_ws, ws_bytes, use_guide, h_mean, h_S, a, seed_u, off_u,
                        seed_n, off_n, scale_u, d_x_new, d_invalid, d_counters);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(n_new >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_draw_move: n_new must be positive");
    cudaStream_t st = as_stream(stream);
    q.n_new = n_new;
    q.postselect = postselect ? 1 : 0;
    QB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, 2 * sizeof(int64_t), st));
    if (!postselect) QB_CUDA_CHECK(cudaMemsetAsync(d_invalid, 0, static_cast<size_t>(n_new), st));
    const int grid = capped_grid(((n_new + 1) / 2 + 255) / 256, 8);
    switch (d) {
        case 1: lw_draw_move_kernel<1><<<grid, 256, 0, st>>>(q); break;
        case 2: lw_draw_move_kernel<2><<<grid, 256, 0, st>>>(q); break;
        case 3: lw_draw_move_kernel<3><<<grid, 256, 0, st>>>(q); break;
        default: lw_draw_move_kernel<4><<<grid, 256, 0, st>>>(q); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_draw_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                                const double* d_cdf, const void* d_ws, size_t ws_bytes, int32_t use_guide,
                                const double* h_mean, const double* h_S, double a, uint64_t seed_u, uint64_t off_u,
                                uint64_t seed_n, uint64_t off_n, int32_t scale_u, const int64_t* d_idxs, int64_t k,
                                int32_t own_mean, double* d_x_new, uint8_t* d_invalid, int64_t* d_counters,
                                void* stream) {
    LwFusedParams q;
    int rc = fill_fused(q, model, d_x_old, n_old, d, d_cdf, d_ws, ws_bytes, use_guide, h_mean, h_S, a, seed_u, off_u,
                        seed_n, off_n, scale_u, d_x_new, d_invalid, d_counters);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_idxs && k >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_draw_retry: bad arguments");
    cudaStream_t st = as_stream(stream);
    q.idxs = d_idxs;
    q.k = k;
    q.own_mean = own_mean ? 1 : 0;
    QB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, sizeof(int64_t), st));  // the clamp count of the first pass stays
    const int grid = capped_grid((k + 127) / 128, 8);
    switch (d) {
        case 1: lw_draw_retry_kernel<1><<<grid, 128, 0, st>>>(q); break;
        case 2: lw_draw_retry_kernel<2><<<grid, 128, 0, st>>>(q); break;
        case 3: lw_draw_retry_kernel<3><<<grid, 128, 0, st>>>(q); break;
        default: lw_draw_retry_kernel<4><<<grid, 128, 0, st>>>(q); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

// =============================================================================================================
// Merge draw (device-RNG mode, d <= 4): the uniforms are generated ALREADY SORTED, so the multinomial draw becomes
// a streaming merge of two ascending sequences instead of n random bisections.
//
//   E_0 .. E_n  i.i.d. Exp(1)  (E_k = -log(1 - philox uniform k)),   S_k = E_0 + .. + E_k,   U_k = S_k / S_n
// are distributed exactly as the order statistics of n i.i.d. U(0,1) variates (exponential spacings), and the new
// particles of a resample are exchangeable, so slot k may take the k-th smallest uniform.  Three launches:
//   exp_tile_sums_kernel   per-tile sums of E            (Philox + log, no memory traffic but the tile sums)
//   cdf_scan_tiles_kernel  exclusive scan of the tile sums (shared with the CDF)
//   lw_merge_move_kernel   per tile: regenerate E, block scan -> 8 consecutive ascending uniforms per thread;
//                          ONE guided bisection for the thread's first uniform, then a forward walk through the
//                          CDF (neighbouring threads walk neighbouring entries: coalesced, every sector fully
//                          used, parents read once); gather, shrink, perturb, validity, 64-byte stores per thread.
// The guided kernel above needs ~4 random 32-byte sectors per draw over 200 MB of tables (1.4 GB of sector traffic
// at n = 1e7; TLB-bound at n = 1e8); this one streams the CDF and the parents once.  Offspring come out ordered by
// parent index.  The retry pass re-centres on the particle's OWN parent (kept for the invalid slots only): the
// reference's prefix-of-the-original-means quirk (resamplers.py:372) would pick the lowest-index parents here.
// =============================================================================================================
struct MergeParams {
    LwFusedParams f;          // x_old, cdf, guide header/table, x_new, invalid, counters, seeds (seed_u/off_u = the
                              // exponential stream), a, S, ms, model, n_old, n_new, scale_u, postselect
    double* tiles;            // [ntiles + 1] exponential tile sums -> exclusive prefix, total last
    int32_t* parent_inv;      // parent index of every slot found invalid (retry input); may be NULL if !postselect
    double* u_out;            // optional: the sorted uniforms (tests)
    int64_t* js_out;          // optional: every slot's parent (tests)
};

// the 8 exponentials of slots base .. base+7 (base even); slots beyond `last` (inclusive bound) give 0
__device__ __forceinline__ void merge_exponentials(const LwFusedParams& f, int64_t base, int64_t last, double (&e)[SCAN_ITEMS]) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 2; ++q) {
        const int64_t k = base + 2 * q;
        double a = 0.0, b = 0.0;
        if (k <= last) {
            philox_uniform_pair(f.seed_u, f.off_u + static_cast<uint64_t>(k >> 1), a, b);
            a = -log(1.0 - a);  // 1 - a in (0, 1]
            b = (k + 1 <= last) ? -log(1.0 - b) : 0.0;
        }
        e[2 * q] = a;
        e[2 * q + 1] = b;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) exp_tile_sums_kernel(const __grid_constant__ LwFusedParams f,
                                                                     double* __restrict__ tile_sums) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const int64_t last = f.n_new;  // slots 0 .. n_new (n_new + 1 exponentials)
    const int64_t ntiles = (f.n_new + 1 + SCAN_TILE - 1) / SCAN_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * SCAN_TILE + static_cast<int64_t>(threadIdx.x) * SCAN_ITEMS;
        double e[SCAN_ITEMS];
        merge_exponentials(f, base, last, e);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) s += e[k];
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

template <int D>
__global__ void __launch_bounds__(SCAN_THREADS) lw_merge_move_kernel(const __grid_constant__ MergeParams mp) {
    __shared__ double warp_tot[SCAN_THREADS / 32];
    const LwFusedParams& p = mp.f;
    const DrawCtx dc = make_draw_ctx(p);
    const int64_t n = p.n_old;
    const int64_t last = p.n_new;
    const int64_t ntiles = (p.n_new + 1 + SCAN_TILE - 1) / SCAN_TILE;
    const double inv_total = 1.0 / mp.tiles[ntiles];
    const int lane = threadIdx.x & 31;
    unsigned int over = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * SCAN_TILE + static_cast<int64_t>(threadIdx.x) * SCAN_ITEMS;
        double e[SCAN_ITEMS];
        merge_exponentials(p, base, last, e);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) s += e[k];
        double total;
        double run = mp.tiles[t] + block_exclusive_scan(s, warp_tot, total);
        unsigned int nbad = 0;
        if (base < p.n_new) {
            // ascending uniforms of my slots (clamped below 1: S_k / S_n < 1 up to rounding)
            double u[SCAN_ITEMS];
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                run += e[k];
                u[k] = fmin(run * inv_total, 0.99999999999999989) * dc.scale;
            }
            // parent of the first slot by guided bisection; the parents of the other slots lie a few entries further
            // on: fetch a window of the CDF with independent loads (one memory round trip instead of a chain of
            // dependent ones) and count, per slot, the entries at or below its uniform; walk on only past the window
            unsigned int dummy = 0;
            DrawCtx d1 = dc;
            d1.scale = 1.0;  // u is already scaled
            const int64_t j0 = guided_draw(d1, u[0], dummy);
            const bool clamped0 = __ldg(p.cdf + j0) <= u[0];  // uniform at or beyond the CDF total (then all are)
            constexpr int WIN = 16;
            double cw[WIN];
#pragma unroll
            for (int i = 0; i < WIN; ++i) cw[i] = (j0 + i < n) ? __ldg(p.cdf + j0 + i) : INFINITY;
            int64_t par[SCAN_ITEMS];
            par[0] = j0;
#pragma unroll
            for (int k = 1; k < SCAN_ITEMS; ++k) {
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < WIN; ++i) cnt += (cw[i] <= u[k]) ? 1 : 0;
                int64_t j = j0 + cnt;
                if (cnt == WIN) {
                    while (j < n && __ldg(p.cdf + j) <= u[k]) ++j;
                }
                par[k] = j;
            }
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (clamped0 || par[k] >= n) {
                    par[k] = n - 1;
                    if (base + k < p.n_new) ++over;
                }
            }
            // normals, move, validity
            double out[SCAN_ITEMS][D];
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 2; ++q) {
                const int64_t i0 = base + 2 * q;
                double ev[2][D];
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    const int64_t f0 = static_cast<int64_t>(m) * p.n_new + i0;
                    if ((f0 & 1) == 0) {
                        philox_normal_pair(p.seed_n, p.off_n + static_cast<uint64_t>(f0 >> 1), ev[0][m], ev[1][m]);
                    } else {
                        ev[0][m] = philox_normal_elem(p.seed_n, p.off_n, f0);
                        ev[1][m] = philox_normal_elem(p.seed_n, p.off_n, f0 + 1);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t src = par[2 * q + h];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        double z = 0.0;
#pragma unroll
                        for (int m = 0; m < D; ++m) z = fma(p.S[c * D + m], ev[h][m], z);
                        out[2 * q + h][c] = ((p.a * __ldg(p.x_old + src * D + c)) + p.ms[c]) + z;
                    }
                }
            }
            unsigned long long flags = 0ull;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                const int64_t i = base + k;
                if (i < p.n_new) {
#pragma unroll
                    for (int c = 0; c < D; ++c) stg_stream(p.x_new + i * D + c, out[k][c]);
                    if (mp.u_out != nullptr) mp.u_out[i] = u[k];
                    if (mp.js_out != nullptr) mp.js_out[i] = par[k];
                    if (p.postselect) {
                        auto row = [&](int c) { return out[k][c]; };
                        if (!model_valid(p.mv, row)) {
                            flags |= 1ull << (8 * k);
                            ++nbad;
                            mp.parent_inv[i] = static_cast<int32_t>(par[k]);
                        }
                    }
                }
            }
            if (p.postselect) {
                if (base + SCAN_ITEMS <= p.n_new) {
                    *reinterpret_cast<unsigned long long*>(p.invalid + base) = flags;  // base % 8 == 0
                } else {
                    for (int k = 0; k < SCAN_ITEMS && base + k < p.n_new; ++k)
                        p.invalid[base + k] = static_cast<uint8_t>((flags >> (8 * k)) & 1ull);
                }
            }
        }
        const unsigned int tot = __reduce_add_sync(0xffffffffu, nbad);
        if (tot && lane == 0) atomicAdd(p.counters, static_cast<unsigned long long>(tot));
    }
    if (over) atomicAdd(p.counters + 1, static_cast<unsigned long long>(over));
}

template <int D>
__global__ void __launch_bounds__(128) lw_merge_retry_kernel(const __grid_constant__ MergeParams mp) {
    const LwFusedParams& p = mp.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.k; r += stride) {
        const int64_t dst = p.idxs[r];
        const int64_t src = mp.parent_inv[dst];
        double out[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double z = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m)
                z = fma(p.S[c * D + m], philox_normal_elem(p.seed_n, p.off_n, static_cast<int64_t>(m) * p.k + r), z);
            out[c] = ((p.a * __ldg(p.x_old + src * D + c)) + p.ms[c]) + z;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) p.x_new[dst * D + c] = out[c];
        auto row = [&](int c) { return out[c]; };
        const bool ok = model_valid(p.mv, row);
        p.invalid[dst] = ok ? 0 : 1;
        if (!ok) atomicAdd(p.counters, 1ull);
    }
}

extern "C" int qb_lw_merge_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                                const double* d_cdf, void* d_ws, size_t ws_bytes, int32_t use_guide,
                                const double* h_mean, const double* h_S, double a, uint64_t seed_e, uint64_t off_e,
                                uint64_t seed_n, uint64_t off_n, int32_t scale_u, int64_t n_new, double* d_x_new,
                                int32_t postselect, uint8_t* d_invalid, int32_t* d_parent_inv, int64_t* d_counters,
                                double* d_u_out, int64_t* d_js_out, void* stream) {
    MergeParams mp;
    int rc = fill_fused(mp.f, model, d_x_old, n_old, d, d_cdf, d_ws, ws_bytes, use_guide, h_mean, h_S, a, seed_e,
                        off_e, seed_n, off_n, scale_u, d_x_new, d_invalid, d_counters);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(n_new >= 1 && n_new <= n_old, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_merge_move: needs 1 <= n_new <= n_old (the exponential tile sums reuse the CDF's tile scratch)");
    QB_REQUIRE(d_ws && ws_bytes >= qb_cdf_workspace_bytes(n_old), QB_ERR_WORKSPACE,
               "qb_lw_merge_move: workspace too small");
    QB_REQUIRE(!postselect || d_parent_inv, QB_ERR_INVALID_ARGUMENT, "qb_lw_merge_move: NULL parent buffer");
    QB_REQUIRE((reinterpret_cast<uintptr_t>(d_invalid) & 7) == 0, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_merge_move: d_invalid must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    mp.f.n_new = n_new;
    mp.f.postselect = postselect ? 1 : 0;
    mp.tiles = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);  // the CDF's tile sums are dead
    mp.parent_inv = d_parent_inv;
    mp.u_out = d_u_out;
    mp.js_out = d_js_out;
    QB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, 2 * sizeof(int64_t), st));
    if (!postselect) QB_CUDA_CHECK(cudaMemsetAsync(d_invalid, 0, static_cast<size_t>(n_new), st));
    const int64_t ntiles = (n_new + 1 + SCAN_TILE - 1) / SCAN_TILE;
    const int grid = capped_grid(ntiles, 8);
    exp_tile_sums_kernel<<<grid, SCAN_THREADS, 0, st>>>(mp.f, mp.tiles);
    QB_CUDA_CHECK(cudaGetLastError());
    cdf_scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(mp.tiles, ntiles);
    QB_CUDA_CHECK(cudaGetLastError());
    switch (d) {
        case 1: lw_merge_move_kernel<1><<<grid, SCAN_THREADS, 0, st>>>(mp); break;
        case 2: lw_merge_move_kernel<2><<<grid, SCAN_THREADS, 0, st>>>(mp); break;
        case 3: lw_merge_move_kernel<3><<<grid, SCAN_THREADS, 0, st>>>(mp); break;
        default: lw_merge_move_kernel<4><<<grid, SCAN_THREADS, 0, st>>>(mp); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_merge_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                                 const double* h_mean, const double* h_S, double a, uint64_t seed_n, uint64_t off_n,
                                 const int64_t* d_idxs, int64_t k, const int32_t* d_parent_inv, double* d_x_new,
                                 uint8_t* d_invalid, int64_t* d_counters, void* stream) {
    MergeParams mp;
    static const double dummy_cdf = 1.0;  // never dereferenced by the retry kernel
    int rc = fill_fused(mp.f, model, d_x_old, n_old, d, &dummy_cdf, nullptr, 0, 0, h_mean, h_S, a, 0, 0, seed_n, off_n,
                        0, d_x_new, d_invalid, d_counters);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_idxs && d_parent_inv && k >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_merge_retry: bad arguments");
    cudaStream_t st = as_stream(stream);
    mp.f.idxs = d_idxs;
    mp.f.k = k;
    mp.tiles = nullptr;
    mp.parent_inv = const_cast<int32_t*>(d_parent_inv);
    mp.u_out = nullptr;
    mp.js_out = nullptr;
    QB_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, sizeof(int64_t), st));
    const int grid = capped_grid((k + 127) / 128, 8);
    switch (d) {
        case 1: lw_merge_retry_kernel<1><<<grid, 128, 0, st>>>(mp); break;
        case 2: lw_merge_retry_kernel<2><<<grid, 128, 0, st>>>(mp); break;
        case 3: lw_merge_retry_kernel<3><<<grid, 128, 0, st>>>(mp); break;
        default: lw_merge_retry_kernel<4><<<grid, 128, 0, st>>>(mp); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
