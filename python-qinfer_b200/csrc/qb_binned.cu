// "Binned" multinomial Liu-West resample for the device-RNG mode, d <= 4 (SURVEY §8 a12-a17;
// distributions.py:337-399, resamplers.py:308-372).
//
// The reference draws n i.i.d. uniforms and bisects the global CDF for each (resamplers.py:319-321): n random
// probes into a table far larger than any cache.  A multinomial draw factorises exactly:
//
//     counts per BIN of 2048 consecutive particles  ~  Multinomial(n_new; bin masses)
//     given its count m_t, the offspring of bin t are m_t i.i.d. draws from the bin's own normalised weights
//
// and the new particles of a resample are exchangeable, so bin t's offspring may occupy the contiguous output
// slots [offs[t], offs[t] + m_t).  Every pass then streams:
//
//   binned_sums_kernel   one read of (w, x): per-bin weight sums AND the weighted moments (Sum w, Sum w x,
//                        Sum w x x^T) of distributions.py:337-399; the last block turns the bin sums into the
//                        bin-level CDF `bounds` and publishes the moments (device buffer + pinned host mirror).
//   binned_count_kernel  no HBM traffic: n_new Philox uniforms are located in `bounds` (shared memory) and
//                        histogrammed -> the multinomial counts; the last block scans them into output offsets
//                        and cuts every bin's output range into segments of <= SEG slots.
//                        (Runs while the host takes the d x d matrix square root of the covariance.)
//   binned_move_kernel   per segment: the bin's weights and rows are loaded ONCE, coalesced, into shared memory,
//                        scanned there into the bin-local CDF, and every output slot of the segment draws a fresh
//                        uniform, bisects the shared-memory CDF (no global probe), takes its parent row from
//                        shared memory, shrinks (resamplers.py:325), perturbs (:332), tests validity (:345-357) and
//                        stores its row and its new weight 1/n (:390-392) with coalesced streaming stores.
//                        Invalid slots are appended to a list (slot, parent) for binned_retry_kernel.
//
// HBM traffic per resampled particle: 8(d+1) [sums+moments] + 8(d+1) [bin load] + 8d + 8 [row + weight out]
// = 8(3d+3) B (48 B for d = 1) against SURVEY §8d's 8(3d+5): the global CDF, the uniforms, the indices and the
// guide table of the other draw paths never exist.  The random probes happen in shared memory.
//
// Statistical contract: the offspring multiset has exactly the law of the reference's (n i.i.d. draws from the
// normalised weights); slot order differs (ordered by parent bin), which the resampler's result does not depend
// on (uniform weights, exchangeable particles).  Philox streams: (seed_u, off_u) element i < n_new locates draw i's
// bin; (seed_v, off_v) element i positions slot i inside its bin; normals as in qb_lw_draw_move.  A retry
// re-centres the slot on its own parent with normals indexed by SLOT (deterministic whatever the list order).
//
// Compiled with --fmad=false (one rounding per reference ufunc in the move; explicit fma() in the reductions).
#include <cstdlib>
#include "qb_models.cuh"
#include "qb_philox.cuh"

namespace qb {

constexpr int BIN = 2048;                         // particles per bin
constexpr int BIN_LOG2 = 11;
constexpr int BIN_THREADS = 256;
constexpr int BIN_ITEMS = BIN / BIN_THREADS;      // 8 consecutive particles per thread
constexpr int SEG = 2 * BIN;                      // output slots per segment (mean offspring per bin: BIN)
constexpr int COUNT_THREADS = 1024;
constexpr int BIN_MAX_GRID = 4096;                // partials rows reserved in the workspace
constexpr int MOM_MAX = 1 + 4 + 10;               // packed moment outputs for d <= 4

struct BinLayout {
    size_t bounds, counts, offs, segs, partials, heap, total;
    int64_t T, max_segs, P;
};

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

static BinLayout bin_layout(int64_t n_old, int64_t n_new) {
    BinLayout L;
    L.T = (n_old + BIN - 1) / BIN;
    L.max_segs = L.T + n_new / SEG + 2;
    L.bounds = 512;                                              // [0,512): tickets, nseg, counters (byte 64), Liu-West constants (byte 128)
    L.counts = L.bounds + align256(static_cast<size_t>(L.T + 1) * 8);
    L.offs = L.counts + align256(static_cast<size_t>(L.T + 2) * 4);
    L.partials = L.offs + align256(static_cast<size_t>(L.T + 1) * 8);
    L.P = 1;
    while (L.P < L.T) L.P *= 2;                                  // leaves of the binomial-splitting tree over the bins
    L.heap = L.partials + align256(static_cast<size_t>(BIN_MAX_GRID) * MOM_MAX * 8);
    L.segs = L.heap + align256(static_cast<size_t>(2 * L.P) * 4);  // last: only its size depends on n_new
    L.total = L.segs + align256(static_cast<size_t>(L.max_segs) * 8);
    return L;
}

__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Exclusive block scan of one value per thread (warp shuffles + one shared array of NW entries); `total` = block sum.
template <typename V, int NW>
__device__ __forceinline__ V block_excl_scan(V v, V* warp_tot, V& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    V inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const V t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    V base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const V t = warp_tot[k];
        if (k < wid) base += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return base + (inc - v);
}

// =============================================================================================================
// pass 1: bin sums + moments
// =============================================================================================================
struct BinSumsParams {
    const double* x;
    const double* w;
    const double* stats;
    int64_t n;
    int32_t T, pad;
    double* bounds;          // [T + 1]: bin sums, then (last block) their exclusive prefix; bounds[T] = total
    unsigned int* counts;    // [T + 2]: zeroed here for pass 2
    double* partials;        // [grid][NOUT]
    double* moments_out;     // device, 1 + d + d*d (may be NULL)
    double* mirror;          // pinned host, 32 doubles: values, [29] covariance flags, [30] sqrtm error, tag at [31]
    double tag;
    unsigned int* ticket;
    double* consts;          // != NULL: the last block also derives the Liu-West constants S (16) and (1-a) mean (4)
    double a, h, zero_cov_comp;
};

// Liu-West constants on the device (d <= 4), restating resamplers.py:266-305 + utils.py:593-607 for one thread:
//   cov = E[x x^T] - mu mu^T (distributions.py:388-389); zero Frobenius norm -> zero_cov_comp * I (flag 1);
//   S = h * sqrtm_psd(cov): symmetric eigendecomposition (cyclic Jacobi), eigenvalues <= 0 clipped, V sqrt(w) V^T;
//   err = || S0 S0 - cov ||_F (the host raises ResamplerError when it is not finite; flag 2: cov itself not finite).
// For d = 1 the arithmetic is the host path's, operation for operation (product, difference, sqrt, product).
template <int D>
__device__ void liu_west_consts(const double* out /* 1 + D + D*D moments */, double a, double h, double zero_cov_comp,
                                double* consts /* S[16], ms[4] */, double& flags, double& err) {
    double mean[D], cov[D][D];
#pragma unroll
    for (int m = 0; m < D; ++m) mean[m] = out[1 + m];
    bool finite = true;
    double fro2 = 0.0;
#pragma unroll
    for (int m = 0; m < D; ++m)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            cov[m][c] = out[1 + D + m * D + c] - mean[m] * mean[c];
            finite = finite && isfinite(cov[m][c]);
            fro2 += cov[m][c] * cov[m][c];
        }
    flags = finite ? 0.0 : 2.0;
    if (finite && fro2 == 0.0) {
        flags = 1.0;
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int c = 0; c < D; ++c) cov[m][c] = (m == c) ? zero_cov_comp : 0.0;
    }
    double S0[D][D];
    if (D == 1) {
        S0[0][0] = (cov[0][0] > 0.0) ? sqrt(cov[0][0]) : 0.0;
    } else {
        double A[D][D], V[D][D];
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                A[m][c] = cov[m][c];
                V[m][c] = (m == c) ? 1.0 : 0.0;
            }
        for (int sweep = 0; sweep < 24; ++sweep) {
            double off = 0.0;
#pragma unroll
            for (int p_ = 0; p_ < D; ++p_)
#pragma unroll
                for (int q_ = p_ + 1; q_ < D; ++q_) off += A[p_][q_] * A[p_][q_];
            if (!(off > 0.0)) break;
#pragma unroll
            for (int p_ = 0; p_ < D; ++p_)
#pragma unroll
                for (int q_ = p_ + 1; q_ < D; ++q_) {
                    const double apq = A[p_][q_];
                    if (apq == 0.0) continue;
                    const double theta = (A[q_][q_] - A[p_][p_]) / (2.0 * apq);
                    const double t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
                    for (int k = 0; k < D; ++k) {   // A <- A J
                        const double akp = A[k][p_], akq = A[k][q_];
                        A[k][p_] = cs * akp - sn * akq;
                        A[k][q_] = sn * akp + cs * akq;
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) {   // A <- J^T A
                        const double apk = A[p_][k], aqk = A[q_][k];
                        A[p_][k] = cs * apk - sn * aqk;
                        A[q_][k] = sn * apk + cs * aqk;
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) {   // V <- V J
                        const double vkp = V[k][p_], vkq = V[k][q_];
                        V[k][p_] = cs * vkp - sn * vkq;
                        V[k][q_] = sn * vkp + cs * vkq;
                    }
                    A[p_][q_] = 0.0;
                    A[q_][p_] = 0.0;
                }
        }
        double rt[D];
#pragma unroll
        for (int k = 0; k < D; ++k) rt[k] = (A[k][k] > 0.0) ? sqrt(A[k][k]) : 0.0;   // w[w <= 0] = 0; sqrt(w)
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) v = fma(V[m][k] * rt[k], V[c][k], v);
                S0[m][c] = v;
            }
    }
    double e2 = 0.0;
#pragma unroll
    for (int m = 0; m < D; ++m)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) v = fma(S0[m][k], S0[k][c], v);
            const double df = v - cov[m][c];
            e2 += df * df;
        }
    err = sqrt(e2);   // (overflows to inf where np.linalg.norm's x.dot(x) does: the reference then raises)
    for (int j = 0; j < 16; ++j) consts[j] = 0.0;
#pragma unroll
    for (int m = 0; m < D; ++m)
#pragma unroll
        for (int c = 0; c < D; ++c) consts[m * D + c] = h * S0[m][c];
    const double oma = 1.0 - a;
    for (int c = 0; c < 4; ++c) consts[16 + c] = (c < D) ? oma * mean[c] : 0.0;   // (1 - a) * mean
}

// Tail of pass 1, shared by both kernels: per-block moment partials, last-block-done ticket, and in the last block the
// bin-level CDF, the finished moments, the Liu-West constants and the host mirror.  NW = warps of the block.
template <int D, int NW>
__device__ __forceinline__ void sums_finish(const BinSumsParams& p, const double (&acc)[1 + D + D * (D + 1) / 2],
                                            double* mred, double* scan_tot, unsigned int* is_last_p) {
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;
    constexpr int NTHREADS = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned int& is_last = *is_last_p;
    // moment partials of this block (fixed order)
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) mred[wid * NOUT + k] = v;
    }
    __syncthreads();
    if (tid < NOUT) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) v += mred[k * NOUT + tid];
        p.partials[static_cast<size_t>(blockIdx.x) * NOUT + tid] = v;
    }
    __threadfence();   // every warp's bin sums (lane 0) and the partials, before this block's ticket
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last block: bin sums -> exclusive prefix (the bin-level CDF) ----
    const int T = p.T;
    const int per = (T + NTHREADS - 1) / NTHREADS;
    const int lo = tid * per, hi = (lo + per < T) ? lo + per : T;
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += __ldcg(p.bounds + i);
    double total;
    double run = block_excl_scan<double, NW>(s, scan_tot, total);
    for (int i = lo; i < hi; ++i) {
        const double v = __ldcg(p.bounds + i);
        p.bounds[i] = run;
        run += v;
    }
    if (tid == 0) {
        p.bounds[T] = total;
        p.counts[T] = 0u;
        p.counts[T + 1] = 0u;
    }
    // ---- moments: one warp per output, lanes stride over the blocks' partials (fixed order) ----
    __shared__ double fin[NOUT];
    for (int o = wid; o < NOUT; o += NW) {
        double v = 0.0;
        for (int b = lane; b < static_cast<int>(gridDim.x); b += 32) v += __ldcg(p.partials + static_cast<size_t>(b) * NOUT + o);
        v = warp_sum(v);
        if (lane == 0) fin[o] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double out[1 + D + D * D];
        out[0] = fin[0];
#pragma unroll
        for (int m = 0; m < D; ++m) out[1 + m] = fin[1 + m];
        int o = 1 + D;
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int c = m; c < D; ++c) {
                out[1 + D + m * D + c] = fin[o];
                out[1 + D + c * D + m] = fin[o];
                ++o;
            }
        if (p.moments_out != nullptr)
            for (int k = 0; k < 1 + D + D * D; ++k) p.moments_out[k] = out[k];
        double flags = 0.0, err = 0.0;
        if (p.consts != nullptr) liu_west_consts<D>(out, p.a, p.h, p.zero_cov_comp, p.consts, flags, err);
        if (p.mirror != nullptr) {
            for (int k = 0; k < 1 + D + D * D; ++k) p.mirror[k] = out[k];
            p.mirror[29] = flags;
            p.mirror[30] = err;
            __threadfence_system();
            *reinterpret_cast<volatile double*>(p.mirror + 31) = p.tag;
        }
        *p.ticket = 0u;
    }
}

template <int D>
__global__ void __launch_bounds__(BIN_THREADS) binned_sums_kernel(const __grid_constant__ BinSumsParams p) {
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;
    constexpr int NW = BIN_THREADS / 32;
    constexpr int CHUNK = 32 * BIN_ITEMS;            // particles one warp covers per step (8 per lane)
    constexpr int NCHUNK = BIN / CHUNK;              // 8 steps per bin
    __shared__ double mred[NW * NOUT];
    __shared__ double scan_tot[NW];
    __shared__ unsigned int is_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double inv = p.stats[QB_STAT_INV_NORM];
    double acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.0;
    // One WARP per bin: no block-wide barrier in the streaming loop, every warp keeps its own loads in flight.
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * NW + wid;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * NW;
    for (int64_t t = gw; t < p.T; t += nwarps) {
        double s = 0.0;
#pragma unroll (D <= 2 ? 4 : 2)
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int64_t base = t * BIN + static_cast<int64_t>(ch) * CHUNK + static_cast<int64_t>(lane) * BIN_ITEMS;
            double wv[BIN_ITEMS], xv[BIN_ITEMS * D];
            if (base + BIN_ITEMS <= p.n) {
#pragma unroll
                for (int k = 0; k < BIN_ITEMS / 2; ++k) {
                    const double2 t2 = ldg_stream2(p.w + base + 2 * k);
                    wv[2 * k] = t2.x;
                    wv[2 * k + 1] = t2.y;
                }
#pragma unroll
                for (int k = 0; k < BIN_ITEMS * D / 2; ++k) {
                    const double2 t2 = ldg_stream2(p.x + base * D + 2 * k);
                    xv[2 * k] = t2.x;
                    xv[2 * k + 1] = t2.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < BIN_ITEMS; ++k) {
                    const bool in = base + k < p.n;
                    wv[k] = in ? p.w[base + k] : 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) xv[k * D + c] = in ? p.x[(base + k) * D + c] : 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < BIN_ITEMS; ++k) {
                const double wi = wv[k] * inv;
                s += wi;
                acc[0] += wi;
                int o = 1 + D;
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    const double wx = wi * xv[k * D + m];
                    acc[1 + m] += wx;
#pragma unroll
                    for (int c = m; c < D; ++c) {
                        acc[o] = fma(wx, xv[k * D + c], acc[o]);
                        ++o;
                    }
                }
            }
        }
        s = warp_sum(s);
        if (lane == 0) {
            p.bounds[t] = s;
            p.counts[t] = 0u;
        }
    }
    sums_finish<D, NW>(p, acc, mred, scan_tot, &is_last);
}

// Pass 1, streaming variant: whole bins (w: 16 KB, x: 16 D KB) are staged global -> shared with 1-D bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a STAGES-deep full/empty mbarrier ring by one producer
// lane, like the fused update kernel; 8 consumer warps read 128-bit conflict-free shared loads of particle PAIRS.
// Each consumer warp owns a fixed 256-particle eighth of every bin; the eighth sums meet in shared memory and the
// last warp to arrive adds them in warp order (deterministic) and writes the bin sum.  The ragged last bin (n not a
// multiple of 2048) is read directly by block 0.
constexpr int SUMS_CONSUMERS = 8;
constexpr int SUMS_THREADS = (SUMS_CONSUMERS + 1) * 32;

template <int D, int STAGES>
__global__ void __launch_bounds__(SUMS_THREADS) binned_sums_tma_kernel(const __grid_constant__ BinSumsParams p) {
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;
    constexpr int NW = SUMS_THREADS / 32;
    constexpr uint32_t W_BYTES = BIN * 8u, X_BYTES = BIN * D * 8u, STAGE_BYTES = W_BYTES + X_BYTES;
    extern __shared__ __align__(128) unsigned char ssm[];
    __shared__ double mred[NW * NOUT];
    __shared__ double scan_tot[NW];
    __shared__ double eighth[STAGES][SUMS_CONSUMERS];
    __shared__ unsigned int is_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // full[s] at bar0 + 8 s, done[s] (every consumer warp has stored its part of the bin sum) at bar0 + 32 + 8 s,
    // empty[s] at bar0 + 64 + 8 s
    const uint32_t bar0 = smem_u32(ssm);
    unsigned char* ring = ssm + 128;
    const uint32_t ring0 = smem_u32(ring);
    const int64_t full_bins = p.n / BIN;
    const int my_bins = (full_bins > static_cast<int64_t>(blockIdx.x))
                            ? static_cast<int>((full_bins - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 32 + 8 * s, SUMS_CONSUMERS);
            mbar_init(bar0 + 64 + 8 * s, SUMS_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    double acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.0;
    const double inv = p.stats[QB_STAT_INV_NORM];

    if (wid == SUMS_CONSUMERS) {
        if (lane == 0) {   // ===== producer lane =====
            const uint64_t pol = l2_policy_evict_first();
            int s = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_bins; ++i) {
                const int64_t first = (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x) * BIN;
                mbar_wait(bar0 + 64 + 8 * s, phase ^ 1u);   // passes at once the first time round the ring
                const uint32_t dst = ring0 + static_cast<uint32_t>(s) * STAGE_BYTES;
                mbar_expect_tx(bar0 + 8 * s, STAGE_BYTES);
                tma_load_1d_hint(dst, p.x + first * D, X_BYTES, bar0 + 8 * s, pol);
                tma_load_1d_hint(dst + X_BYTES, p.w + first, W_BYTES, bar0 + 8 * s, pol);
                if (++s == STAGES) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ===== consumer warps =====
        int s = 0;
        uint32_t phase = 0;
        for (int i = 0; i < my_bins; ++i) {
            const int64_t t = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
            const double* xs = reinterpret_cast<const double*>(ring + static_cast<size_t>(s) * STAGE_BYTES);
            const double* ws = reinterpret_cast<const double*>(ring + static_cast<size_t>(s) * STAGE_BYTES + X_BYTES);
            mbar_wait(bar0 + 8 * s, phase);
            double sub = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 256 * wid + 64 * k + 2 * lane;       // my pair of this quarter of the warp's eighth
                const double2 wp = *reinterpret_cast<const double2*>(ws + j);
                double xr[2 * D];
#pragma unroll
                for (int v = 0; v < D; ++v) {
                    const double2 t2 = *reinterpret_cast<const double2*>(xs + static_cast<size_t>(j) * D + 2 * v);
                    xr[2 * v] = t2.x;
                    xr[2 * v + 1] = t2.y;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double wi = (h == 0 ? wp.x : wp.y) * inv;
                    sub += wi;
                    acc[0] += wi;
                    int o = 1 + D;
#pragma unroll
                    for (int m = 0; m < D; ++m) {
                        const double wx = wi * xr[h * D + m];
                        acc[1 + m] += wx;
#pragma unroll
                        for (int c = m; c < D; ++c) {
                            acc[o] = fma(wx, xr[h * D + c], acc[o]);
                            ++o;
                        }
                    }
                }
            }
            sub = warp_sum(sub);
            if (lane == 0) {
                eighth[s][wid] = sub;
                mbar_arrive(bar0 + 32 + 8 * s);               // (release) my part of this bin's sum is in place
                if (wid == 0) {                               // warp 0 collects: fixed-order sum of the eight parts
                    mbar_wait(bar0 + 32 + 8 * s, phase);      // (acquire) every consumer warp has stored its part
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < SUMS_CONSUMERS; ++q) tot += eighth[s][q];
                    p.bounds[t] = tot;
                    p.counts[t] = 0u;
                }
                mbar_arrive(bar0 + 64 + 8 * s);   // only now may the producer refill stage s (and eighth[s] be reused)
            }
            __syncwarp();
            if (++s == STAGES) {
                s = 0;
                phase ^= 1u;
            }
        }
    }
    __syncthreads();
    // ragged last bin: straight from global memory, by block 0 (its sum is written before the ticket below)
    if (blockIdx.x == 0 && full_bins < p.T) {
        const int64_t first = full_bins * BIN;
        const int cnt = static_cast<int>(p.n - first);
        double sub = 0.0;
        for (int j = tid; j < cnt; j += SUMS_THREADS) {
            const double wi = p.w[first + j] * inv;
            sub += wi;
            acc[0] += wi;
            int o = 1 + D;
#pragma unroll
            for (int m = 0; m < D; ++m) {
                const double wx = wi * p.x[(first + j) * D + m];
                acc[1 + m] += wx;
#pragma unroll
                for (int c = m; c < D; ++c) {
                    acc[o] = fma(wx, p.x[(first + j) * D + c], acc[o]);
                    ++o;
                }
            }
        }
        sub = warp_sum(sub);
        if (lane == 0) scan_tot[wid] = sub;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int q = 0; q < NW; ++q) tot += scan_tot[q];
            p.bounds[full_bins] = tot;
            p.counts[full_bins] = 0u;
        }
        __syncthreads();
    }
    sums_finish<D, NW>(p, acc, mred, scan_tot, &is_last);
}

// =============================================================================================================
// pass 2: multinomial counts per bin (histogram of n_new uniforms over the bin-level CDF), offsets, segments
// =============================================================================================================
struct BinCountParams {
    const double* bounds;
    unsigned int* counts;    // [T + 2]
    int64_t* offs;           // [T + 1]
    uint2* segs;             // [max_segs]: (bin, chunk)
    unsigned int* ticket;
    unsigned int* nseg;
    int64_t n_new;
    int32_t T, max_segs;
    uint64_t seed_u, off_u;
};

// bin of v: the t in [0, T) with bounds[t] <= v < bounds[t+1] (clamped).  Starts from the proportional guess
// (bin masses are nearly equal for a well-mixed cloud: 1-3 probes), gallops, then bisects: O(log T) whatever the
// weights look like.
template <typename Tab>
__device__ __forceinline__ int locate_bin(Tab tab, int T, double v, double guess_scale) {
    int g = static_cast<int>(v * guess_scale);
    g = (g < 0) ? 0 : ((g > T - 1) ? T - 1 : g);
    int lo, hi;  // invariant: (lo == 0 or tab(lo) <= v) and (hi == T or tab(hi) > v), lo < hi
    if (tab(g) <= v) {
        lo = g;
        hi = g + 1;
        int step = 1;
        while (hi < T && tab(hi) <= v) {
            lo = hi;
            hi = (hi + step < T) ? hi + step : T;
            step <<= 1;
        }
    } else {
        hi = g;
        lo = g - 1;
        int step = 1;
        while (lo > 0 && tab(lo) > v) {
            hi = lo;
            lo = (lo - step > 0) ? lo - step : 0;
            step <<= 1;
        }
        if (lo < 0) lo = 0;
        if (hi <= lo) hi = lo + 1;
    }
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab(mid) <= v)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

template <bool SMEM>
__global__ void __launch_bounds__(COUNT_THREADS) binned_count_kernel(const __grid_constant__ BinCountParams p) {
    extern __shared__ __align__(16) unsigned char bsm[];
    __shared__ long long scan_a[COUNT_THREADS / 32], scan_b[COUNT_THREADS / 32];
    __shared__ unsigned int is_last;
    const int T = p.T, tid = threadIdx.x;
    double* bounds_s = reinterpret_cast<double*>(bsm);                          // [T + 1]
    unsigned int* hist = reinterpret_cast<unsigned int*>(bounds_s + (T + 2));   // [T + 2] (8-byte aligned, even length)
    const int Tpad = (T + 2) & ~1;
    if (SMEM) {
        for (int i = tid; i <= T; i += COUNT_THREADS) bounds_s[i] = p.bounds[i];
        for (int i = tid; i < Tpad; i += COUNT_THREADS) hist[i] = 0u;
        __syncthreads();
    }
    const double total = SMEM ? bounds_s[T] : p.bounds[T];
    const double guess_scale = static_cast<double>(T) / total;
    auto tab_s = [&](int i) { return bounds_s[i]; };
    auto tab_g = [&](int i) { return __ldg(p.bounds + i); };
    const int64_t npairs = (p.n_new + 1) / 2;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * COUNT_THREADS;
    constexpr int NP = 4;  // Philox counters per thread and iteration: 8 independent searches in flight
    for (int64_t pr0 = static_cast<int64_t>(blockIdx.x) * COUNT_THREADS + tid; pr0 < npairs; pr0 += NP * stride) {
        double v[2 * NP];
        bool live[2 * NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int64_t pr = pr0 + q * stride;
            double u0 = 0.0, u1 = 0.0;
            if (pr < npairs) philox_uniform_pair(p.seed_u, p.off_u + static_cast<uint64_t>(pr), u0, u1);
            v[2 * q] = u0 * total;
            v[2 * q + 1] = u1 * total;
            live[2 * q] = pr < npairs;
            live[2 * q + 1] = pr < npairs && 2 * pr + 1 < p.n_new;
        }
        int b[2 * NP];
#pragma unroll
        for (int q = 0; q < 2 * NP; ++q)
            b[q] = SMEM ? locate_bin(tab_s, T, v[q], guess_scale) : locate_bin(tab_g, T, v[q], guess_scale);
#pragma unroll
        for (int q = 0; q < 2 * NP; ++q)
            if (live[q]) {
                if (SMEM)
                    atomicAdd(hist + b[q], 1u);
                else
                    atomicAdd(p.counts + b[q], 1u);
            }
    }
    if (SMEM) {
        __syncthreads();
        // two 32-bit counts per 64-bit atomic (no carry: every count < n_new < 2^31)
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.counts);
        for (int i = tid; i < Tpad / 2; i += COUNT_THREADS) {
            const unsigned long long v = static_cast<unsigned long long>(hist[2 * i]) |
                                         (static_cast<unsigned long long>(hist[2 * i + 1]) << 32);
            if (v) atomicAdd(dst + i, v);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last block: counts -> output offsets; every bin's output range cut into segments of <= SEG slots ----
    const int per = (T + COUNT_THREADS - 1) / COUNT_THREADS;
    const int lo = tid * per, hi = (lo + per < T) ? lo + per : T;
    long long m_sum = 0, s_sum = 0;
    for (int i = lo; i < hi; ++i) {
        const long long m = static_cast<long long>(__ldcg(p.counts + i));
        m_sum += m;
        s_sum += (m + SEG - 1) / SEG;
    }
    long long m_tot, s_tot;
    long long m_run = block_excl_scan<long long, COUNT_THREADS / 32>(m_sum, scan_a, m_tot);
    long long s_run = block_excl_scan<long long, COUNT_THREADS / 32>(s_sum, scan_b, s_tot);
    for (int i = lo; i < hi; ++i) {
        const long long m = static_cast<long long>(__ldcg(p.counts + i));
        p.offs[i] = m_run;
        m_run += m;
        const long long ns = (m + SEG - 1) / SEG;
        for (long long c = 0; c < ns; ++c)
            if (s_run + c < p.max_segs) p.segs[s_run + c] = make_uint2(static_cast<unsigned int>(i), static_cast<unsigned int>(c));
        s_run += ns;
    }
    if (tid == 0) {
        p.offs[T] = m_tot;
        *p.nseg = static_cast<unsigned int>(s_tot < p.max_segs ? s_tot : p.max_segs);
        // the move / retry kernels' counters: [0] invalid slots (list length), [1] clamped draws, [2] still invalid
        // after a retry launch, [3] most rounds used
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(p.ticket - 1 + 16);
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0ull;
        *p.ticket = 0u;
    }
}

// -------------------------------------------------------------------------------------------------------------
// pass 2, default: the same multinomial counts WITHOUT touching n_new uniforms.  A multinomial over the bins factorises
// over a binary tree: the root holds n_new; a node holding c offspring hands Binomial(c, mass(left) / mass(node)) to
// its left child and the rest to the right one; the leaves (bins) then hold exactly Multinomial(n_new; bin masses).
// One binomial variate per tree node (<= 2T of them) replaces n_new uniform draws + searches + atomics: 60 us -> ~15 us
// at n = 1e7 and 1.3 ms -> ~0.05 ms at n = 1e8, where the histogram's tables no longer fit in shared memory.
// The binomial sampler is exact (inversion for small means, BTPE rejection [Kachitvichyanukul & Schmeiser 1988] with
// its Stirling squeeze otherwise — the algorithm NumPy's legacy generator uses), driven by counter-based Philox
// uniforms: node k consumes counters off_u + 32 k .. of stream seed_u, so the counts are a pure function of
// (weights, seed, offset).  One CTA walks the tree level by level.
// -------------------------------------------------------------------------------------------------------------
struct NodeRng {
    uint64_t seed, ctr;
    double spare;
    int have;
    __device__ __forceinline__ double next() {
        if (have) {
            have = 0;
            return spare;
        }
        double a, b;
        philox_uniform_pair(seed, ctr++, a, b);
        spare = b;
        have = 1;
        return a;
    }
};

__device__ long long binomial_inversion(NodeRng& g, long long n, double p) {   // p <= 1/2, n p < 30
    const double q = 1.0 - p, qn = exp(static_cast<double>(n) * log(q)), np = static_cast<double>(n) * p;
    const double bd = np + 10.0 * sqrt(np * q + 1.0);
    const long long bound = (static_cast<double>(n) < bd) ? n : static_cast<long long>(bd);
    long long X = 0;
    double px = qn, U = g.next();
    while (U > px) {
        ++X;
        if (X > bound) {
            X = 0;
            px = qn;
            U = g.next();
        } else {
            U -= px;
            px = ((static_cast<double>(n - X + 1)) * p * px) / (static_cast<double>(X) * q);
        }
    }
    return X;
}

__device__ __forceinline__ double stirling_tail(double v, double v2) {
    return (13680.0 - (462.0 - (132.0 - (99.0 - 140.0 / v2) / v2) / v2) / v2) / v / 166320.0;
}

__device__ long long binomial_btpe(NodeRng& g, long long n_, double p) {        // p <= 1/2, n p >= 30
    const double n = static_cast<double>(n_);
    const double r = p, q = 1.0 - r, fm = n * r + r;
    const double m = floor(fm);
    const double p1 = floor(2.195 * sqrt(n * r * q) - 4.6 * q) + 0.5;
    const double xm = m + 0.5, xl = xm - p1, xr = xm + p1;
    const double c = 0.134 + 20.5 / (15.3 + m);
    double a = (fm - xl) / (fm - xl * r);
    const double laml = a * (1.0 + a / 2.0);
    a = (xr - fm) / (xr * q);
    const double lamr = a * (1.0 + a / 2.0);
    const double p2 = p1 * (1.0 + 2.0 * c), p3 = p2 + c / laml, p4 = p3 + c / lamr;
    const double nrq = n * r * q;
    double y;
    for (;;) {
        const double u = g.next() * p4;
        double v = g.next();
        if (u <= p1) {                                   // triangular centre: accept at once
            y = floor(xm - p1 * v + u);
            break;
        }
        if (u <= p2) {                                   // parallelograms
            const double x = xl + (u - p1) / c;
            v = v * c + 1.0 - fabs(m - x + 0.5) / p1;
            if (v > 1.0) continue;
            y = floor(x);
        } else if (u <= p3) {                            // left exponential tail
            y = floor(xl + log(v) / laml);
            if (y < 0.0 || v == 0.0) continue;
            v = v * (u - p2) * laml;
        } else {                                         // right exponential tail
            y = floor(xr - log(v) / lamr);
            if (y > n || v == 0.0) continue;
            v = v * (u - p3) * lamr;
        }
        const double k = fabs(y - m);
        if (!(k > 20.0 && k < nrq / 2.0 - 1.0)) {        // explicit evaluation of f(y) / f(m)
            const double s = r / q, aa = s * (n + 1.0);
            double F = 1.0;
            if (m < y) {
                for (double i = m + 1.0; i <= y; i += 1.0) F *= (aa / i - s);
            } else if (m > y) {
                for (double i = y + 1.0; i <= m; i += 1.0) F /= (aa / i - s);
            }
            if (v > F) continue;
            break;
        }
        // squeezing with the Stirling series
        const double rho = (k / nrq) * ((k * (k / 3.0 + 0.625) + 0.16666666666666666) / nrq + 0.5);
        const double t = -k * k / (2.0 * nrq);
        const double A = log(v);
        if (A < t - rho) break;
        if (A > t + rho) continue;
        const double x1 = y + 1.0, f1 = m + 1.0, z = n + 1.0 - m, w = n - y + 1.0;
        const double bound = xm * log(f1 / x1) + (n - m + 0.5) * log(z / w) + (y - m) * log(w * r / (x1 * q)) +
                             stirling_tail(f1, f1 * f1) + stirling_tail(z, z * z) + stirling_tail(x1, x1 * x1) +
                             stirling_tail(w, w * w);
        if (A > bound) continue;
        break;
    }
    return static_cast<long long>(y);
}

// X ~ Binomial(n, p), exact
__device__ long long binomial_sample(uint64_t seed, uint64_t ctr, long long n, double p) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    NodeRng g = {seed, ctr, 0.0, 0};
    const double r = (p <= 0.5) ? p : 1.0 - p;
    const long long y = (static_cast<double>(n) * r < 30.0) ? binomial_inversion(g, n, r) : binomial_btpe(g, n, r);
    return (p <= 0.5) ? y : n - y;
}

__global__ void __launch_bounds__(256) binomial_test_kernel(uint64_t seed, uint64_t off, long long n, double p,
                                                            int64_t count, long long* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = binomial_sample(seed, off + 32ull * static_cast<uint64_t>(i), n, p);
}

struct TreeParams {
    BinCountParams c;
    unsigned int* heap;   // [2P]: node k (1-based, children 2k and 2k+1) holds its offspring count; leaves at P + t
    int32_t P, levels;
};

// one node of the splitting tree: hand Binomial(c, mass(left) / mass(node)) to the left child
__device__ __forceinline__ void split_node(const TreeParams& tp, int k, int lo, int span) {
    const BinCountParams& p = tp.c;
    const int T = p.T;
    const long long c = static_cast<long long>(tp.heap[k]);
    const int mid = lo + span / 2, hi = lo + span;
    long long left = 0;
    if (c > 0) {
        if (mid >= T) {
            left = c;                                   // the right half lies beyond the last bin
        } else {
            const double b_lo = p.bounds[lo], b_mid = p.bounds[mid], b_hi = p.bounds[(hi < T) ? hi : T];
            const double mass = b_hi - b_lo;
            const double ql = (mass > 0.0) ? (b_mid - b_lo) / mass : 1.0;
            left = binomial_sample(p.seed_u, p.off_u + 32ull * static_cast<uint64_t>(k), c,
                                   (ql < 0.0) ? 0.0 : ((ql > 1.0) ? 1.0 : ql));
        }
    }
    tp.heap[2 * k] = static_cast<unsigned int>(left);
    tp.heap[2 * k + 1] = static_cast<unsigned int>(c - left);
}

// leaf counts -> output offsets; every bin's output range cut into segments of <= SEG slots (one CTA)
__device__ void tree_leaf_scan(const TreeParams& tp, long long* scan_a, long long* scan_b) {
    const BinCountParams& p = tp.c;
    const int T = p.T, tid = threadIdx.x;
    const unsigned int* leaf = tp.heap + tp.P;
    const int per = (T + COUNT_THREADS - 1) / COUNT_THREADS;
    const int lo = tid * per, hi = (lo + per < T) ? lo + per : T;
    long long m_sum = 0, s_sum = 0;
    for (int i = lo; i < hi; ++i) {
        const long long m = static_cast<long long>(__ldcg(leaf + i));
        m_sum += m;
        s_sum += (m + SEG - 1) / SEG;
    }
    long long m_tot, s_tot;
    long long m_run = block_excl_scan<long long, COUNT_THREADS / 32>(m_sum, scan_a, m_tot);
    long long s_run = block_excl_scan<long long, COUNT_THREADS / 32>(s_sum, scan_b, s_tot);
    for (int i = lo; i < hi; ++i) {
        const long long m = static_cast<long long>(__ldcg(leaf + i));
        p.counts[i] = static_cast<unsigned int>(m);
        p.offs[i] = m_run;
        m_run += m;
        const long long ns = (m + SEG - 1) / SEG;
        for (long long c = 0; c < ns; ++c)
            if (s_run + c < p.max_segs) p.segs[s_run + c] = make_uint2(static_cast<unsigned int>(i), static_cast<unsigned int>(c));
        s_run += ns;
    }
    if (tid == 0) {
        p.offs[T] = m_tot;
        *p.nseg = static_cast<unsigned int>(s_tot < p.max_segs ? s_tot : p.max_segs);
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(p.ticket - 1 + 16);
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0ull;
    }
}

// levels [0, top) of the tree by one CTA (the upper levels have few nodes: their cost is the dependency chain); when
// that is the whole tree, the leaf scan follows in the same launch
__global__ void __launch_bounds__(COUNT_THREADS) binned_tree_count_kernel(const __grid_constant__ TreeParams tp, int top) {
    __shared__ long long scan_a[COUNT_THREADS / 32], scan_b[COUNT_THREADS / 32];
    const int P = tp.P, tid = threadIdx.x;
    if (tid == 0) tp.heap[1] = static_cast<unsigned int>(tp.c.n_new);
    __syncthreads();
    for (int lvl = 0; lvl < top; ++lvl) {
        const int nodes = 1 << lvl;
        const int span = P >> lvl;                              // leaves under a node of this level
        for (int i = tid; i < nodes; i += COUNT_THREADS) split_node(tp, nodes + i, i * span, span);
        __syncthreads();
    }
    if (top == tp.levels) tree_leaf_scan(tp, scan_a, scan_b);
}

// levels [top, levels): one CTA per node of level `top`, each finishing its own subtree
__global__ void __launch_bounds__(64) binned_subtree_kernel(const __grid_constant__ TreeParams tp, int top) {
    const int P = tp.P;
    const int root = blockIdx.x;                                // index within level `top`
    for (int lvl = top; lvl < tp.levels; ++lvl) {
        const int rel = 1 << (lvl - top);                       // nodes of this level under my root
        const int span = P >> lvl;
        for (int j = threadIdx.x; j < rel; j += blockDim.x) {
            const int i = root * rel + j;
            split_node(tp, (1 << lvl) + i, i * span, span);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(COUNT_THREADS) binned_leaf_scan_kernel(const __grid_constant__ TreeParams tp) {
    __shared__ long long scan_a[COUNT_THREADS / 32], scan_b[COUNT_THREADS / 32];
    tree_leaf_scan(tp, scan_a, scan_b);
}

// =============================================================================================================
// pass 3: per segment, bin-local CDF in shared memory, draw + gather + shrink + perturb + validity + weights
// =============================================================================================================
struct BinMoveParams {
    const double* x_old;
    const double* w;
    const double* stats;
    const int64_t* offs;
    const uint2* segs;
    const unsigned int* nseg;
    double* x_new;
    double* x_new2;          // slots >= split are stored at x_new2[(slot - split)] (sharded: surplus rows); may be NULL
    int64_t split;
    double* w_new;           // may be NULL
    double w_value;
    double* stats_new;       // may be NULL: stats block of the uniform weights
    double n_global;
    long long* list;         // invalid slots: slot | parent << 32 (postselect only)
    unsigned long long* counters;  // [0] invalid, [1] clamped draws
    int64_t* js_out;         // optional: global parent index of every slot (tests)
    int32_t* parents;        // global parent index of every slot (postselect): a uniformly random slot's parent is an
                             // i.i.d. draw from the weighted cloud — what the reference's retry re-centres on
    int32_t own_mean, pad1;  // retry: 1 = the slot's own parent, 0 = the reference's law (resamplers.py:372, below)
    double* mirror;          // pinned host: {invalid, clamped, drawn, tag} (may be NULL)
    double tag;
    unsigned int* ticket;
    int64_t n_old, n_new;
    uint64_t round_stride;   // retry: Philox counters between the normal streams of consecutive rounds
    int32_t T, postselect, rounds, pad0;
    uint64_t seed_v, off_v, seed_n, off_n;
    double a;
    double S[16];
    double ms[4];
    const double* consts;    // != NULL: S and (1-a) mean come from the workspace (derived on the device by pass 1)
    ModelView mv;
};

// the Liu-West constants of this launch: launch parameters (host-derived) or the workspace (device-derived)
template <int D>
__device__ __forceinline__ void load_consts(const BinMoveParams& p, double (&S)[D * D], double (&ms)[D]) {
    if (p.consts != nullptr) {
#pragma unroll
        for (int j = 0; j < D * D; ++j) S[j] = __ldcg(p.consts + j);
#pragma unroll
        for (int c = 0; c < D; ++c) ms[c] = __ldcg(p.consts + 16 + c);
    } else {
#pragma unroll
        for (int j = 0; j < D * D; ++j) S[j] = p.S[j];
#pragma unroll
        for (int c = 0; c < D; ++c) ms[c] = p.ms[c];
    }
}

template <int D>
__device__ __forceinline__ double* slot_ptr(const BinMoveParams& p, int64_t i) {
    return (p.x_new2 != nullptr && i >= p.split) ? p.x_new2 + (i - p.split) * D : p.x_new + i * D;
}

// normals eps[m] of slot pair (i0, i0 + 1), i0 even: element m * n_new + i of stream (seed_n, off_n)
template <int D>
__device__ __forceinline__ void slot_normals(const BinMoveParams& p, int64_t i0, bool need0, bool need1,
                                             double (&ev)[2][D]) {
#pragma unroll
    for (int m = 0; m < D; ++m) {
        const int64_t f0 = static_cast<int64_t>(m) * p.n_new + i0;
        if ((f0 & 1) == 0) {
            philox_normal_pair(p.seed_n, p.off_n + static_cast<uint64_t>(f0 >> 1), ev[0][m], ev[1][m]);
        } else {
            ev[0][m] = need0 ? philox_normal_elem(p.seed_n, p.off_n, f0) : 0.0;
            ev[1][m] = need1 ? philox_normal_elem(p.seed_n, p.off_n, f0 + 1) : 0.0;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(BIN_THREADS, (D <= 2) ? 3 : 2) binned_move_kernel(const __grid_constant__ BinMoveParams p) {
    extern __shared__ __align__(16) unsigned char msm[];
    constexpr int NW = BIN_THREADS / 32;
    double* cdf_s = reinterpret_cast<double*>(msm);        // [BIN]
    double* x_s = cdf_s + BIN;                             // [BIN * D]
    __shared__ double scan_tot[NW];
    __shared__ unsigned int is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const double inv = p.stats[QB_STAT_INV_NORM];
    const unsigned int nseg = *p.nseg;
    unsigned int n_clamped = 0;
    double Sl[D * D], msl[D];
    load_consts<D>(p, Sl, msl);
    if (blockIdx.x == 0 && tid == 0 && p.stats_new != nullptr) {
        // resamplers.py:390-392: weights 1/n; the stats block of set_uniform_kernel
        double* st = p.stats_new;
        st[QB_STAT_NORM] = 1.0;
        st[QB_STAT_SUMSQ] = p.w_value;
        st[QB_STAT_MIN] = p.w_value;
        st[QB_STAT_NBAD] = 0.0;
        st[QB_STAT_INV_NORM] = 1.0;
        st[QB_STAT_NESS] = p.n_global;
        st[QB_STAT_TAG] = 0.0;
        st[QB_STAT_SKIPPED] = 0.0;
        st[QB_STAT_ATTN] = 0.0;
    }
    for (unsigned int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const uint2 sg = p.segs[s];
        const int64_t bin = sg.x;
        const int64_t first = bin * BIN;
        const int cnt = static_cast<int>((p.n_old - first < BIN) ? (p.n_old - first) : BIN);
        const int64_t o_bin = p.offs[bin];
        const int64_t o_begin = o_bin + static_cast<int64_t>(sg.y) * SEG;
        int64_t o_end = p.offs[bin + 1];
        if (o_end > o_begin + SEG) o_end = o_begin + SEG;
        // ---- load the bin, scan the weights into the bin-local CDF ----
        const int j0 = tid * BIN_ITEMS;
        double wv[BIN_ITEMS];
        if (cnt == BIN) {
#pragma unroll
            for (int k = 0; k < BIN_ITEMS / 2; ++k) {
                const double2 t2 = *reinterpret_cast<const double2*>(p.w + first + j0 + 2 * k);
                wv[2 * k] = t2.x * inv;
                wv[2 * k + 1] = t2.y * inv;
            }
#pragma unroll
            for (int k = 0; k < BIN_ITEMS * D / 2; ++k) {
                const double2 t2 = *reinterpret_cast<const double2*>(p.x_old + (first + j0) * D + 2 * k);
                *reinterpret_cast<double2*>(x_s + j0 * D + 2 * k) = t2;
            }
        } else {
#pragma unroll
            for (int k = 0; k < BIN_ITEMS; ++k) {
                const bool in = j0 + k < cnt;
                wv[k] = in ? p.w[first + j0 + k] * inv : 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) x_s[(j0 + k) * D + c] = in ? p.x_old[(first + j0 + k) * D + c] : 0.0;
            }
        }
        double ts = 0.0;
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; ++k) ts += wv[k];
        double total;
        double run = block_excl_scan<double, NW>(ts, scan_tot, total);
        // The bin-local CDF goes to shared memory in EYTZINGER (breadth-first) order: node k of the implicit search
        // tree over cdf[0 .. BIN-2] sits at eyt[k], children at 2k and 2k+1.  The probes of one tree level are
        // contiguous, so the upper levels (few distinct nodes per warp) are conflict-free — in sorted order they sit
        // 2^j entries apart, i.e. in ONE bank.  (Entry BIN-1, the bin total, is never needed: v < total.)
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; ++k) {
            run += wv[k];
            const unsigned int r = static_cast<unsigned int>(j0 + k) + 1u;   // 1-based sorted rank
            if (r < static_cast<unsigned int>(BIN)) {
                const int tz = __ffs(r) - 1;
                cdf_s[(1u << (BIN_LOG2 - 1 - tz)) + (r >> (tz + 1))] = run;
            } else {
                cdf_s[0] = run;                                               // slot 0 is no tree node: keeps the total
            }
        }
        __syncthreads();
        const double top = cdf_s[0];
        // ---- the segment's output slots: NP Philox counters (2 slots each) per thread and round ----
        constexpr int NP = 4;
        const int64_t base_even = o_begin & ~1LL;
        for (int64_t ib = base_even + 2 * tid; ib < o_end; ib += 2 * NP * BIN_THREADS) {
            double v[2 * NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int64_t i0 = ib + static_cast<int64_t>(q) * 2 * BIN_THREADS;
                double u0 = 0.0, u1 = 0.0;
                if (i0 < o_end) philox_uniform_pair(p.seed_v, p.off_v + static_cast<uint64_t>(i0 >> 1), u0, u1);
                v[2 * q] = u0 * top;
                v[2 * q + 1] = u1 * top;
            }
            // branch-free descent, the 2 NP searches interleaved: k <- 2k + (eyt[k] <= v); rank = k - BIN = number
            // of entries <= v = first index with cdf > v (side='right', resamplers.py:321)
            unsigned int kk[2 * NP];
#pragma unroll
            for (int q = 0; q < 2 * NP; ++q) kk[q] = 1u;
#pragma unroll
            for (int lvl = 0; lvl < BIN_LOG2; ++lvl) {
#pragma unroll
                for (int q = 0; q < 2 * NP; ++q) kk[q] = 2u * kk[q] + ((cdf_s[kk[q]] <= v[q]) ? 1u : 0u);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int64_t i0 = ib + static_cast<int64_t>(q) * 2 * BIN_THREADS;
                if (i0 >= o_end) continue;
                const bool need0 = i0 >= o_begin, need1 = i0 + 1 < o_end;
                int par[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int lo = static_cast<int>(kk[2 * q + h]) - BIN;
                    if (lo >= cnt) {                                          // distributions.py:330-333's clamp
                        lo = cnt - 1;
                        if (h == 0 ? need0 : need1) ++n_clamped;
                    }
                    par[h] = lo;
                }
                double ev[2][D];
                slot_normals<D>(p, i0, need0, need1, ev);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!(h == 0 ? need0 : need1)) continue;
                    const int64_t i = i0 + h;
                    double out[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        double z = 0.0;
#pragma unroll
                        for (int m = 0; m < D; ++m) z = fma(Sl[c * D + m], ev[h][m], z);
                        out[c] = ((p.a * x_s[par[h] * D + c]) + msl[c]) + z;  // resamplers.py:325,332, one rounding per ufunc
                    }
                    double* dst = slot_ptr<D>(p, i);
#pragma unroll
                    for (int c = 0; c < D; ++c) stg_stream(dst + c, out[c]);
                    if (p.w_new != nullptr) stg_stream(p.w_new + i, p.w_value);
                    if (p.js_out != nullptr) p.js_out[i] = first + par[h];
                    if (p.parents != nullptr) p.parents[i] = static_cast<int32_t>(first + par[h]);
                    if (p.postselect) {
                        auto row = [&](int c) { return out[c]; };
                        if (!model_valid(p.mv, row)) {
                            const unsigned long long pos = atomicAdd(p.counters, 1ull);
                            p.list[pos] = static_cast<long long>(i) | (static_cast<long long>(first + par[h]) << 32);
                        }
                    }
                }
            }
        }
        __syncthreads();  // the next segment overwrites the shared-memory bin
    }
    n_clamped = __reduce_add_sync(0xffffffffu, n_clamped);
    if (n_clamped && lane == 0) atomicAdd(p.counters + 1, static_cast<unsigned long long>(n_clamped));
    if (p.mirror == nullptr) return;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
        if (is_last) {
            __threadfence();
            const double bad = static_cast<double>(__ldcg(p.counters));
            const double clamped = static_cast<double>(__ldcg(p.counters + 1));
            const double drawn = static_cast<double>(p.offs[p.T]);
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror), "d"(bad), "d"(clamped), "d"(drawn),
                         "d"(p.tag)
                         : "memory");
            *p.ticket = 0u;
        }
    }
}

// Retry over the entries of the invalid list (resamplers.py:327-372).  Every entry takes up to `rounds` fresh
// perturbations in THIS launch — round j uses normals eps[m] = element m * n_new + slot of stream
// (seed_n, off_n + j * round_stride) — and stops at the first valid one; an entry whose slot became valid is marked
// resolved (-1).  The list length is read from the device (counters[0], written by the move kernel), so the host
// queues this launch right behind the move without a round trip; with no invalid slot it exits at once.
// counters[2] = entries still invalid, counters[3] = most rounds any entry used.
template <int D>
__global__ void __launch_bounds__(128) binned_retry_kernel(const __grid_constant__ BinMoveParams p) {
    __shared__ unsigned int is_last;
    const int64_t k0 = static_cast<int64_t>(*reinterpret_cast<const volatile unsigned long long*>(p.counters));
    double Sl[D * D], msl[D];
    load_consts<D>(p, Sl, msl);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < k0; r += stride) {
        const long long e = p.list[r];
        if (e < 0) continue;
        const int64_t slot = e & 0xffffffffLL;
        int64_t parent = e >> 32;
        double mu[D], out[D];
        bool ok = false;
        int used = 0;
        for (int j = 0; j < p.rounds && !ok; ++j) {
            const uint64_t off = p.off_n + static_cast<uint64_t>(j) * p.round_stride;
            if (!p.own_mean && p.parents != nullptr) {
                // The reference re-centres the r-th still-invalid particle on mus[r], the shrunk mean of the r-th
                // ORIGINAL draw (resamplers.py:372 re-slices `mus = mus[:k]`): an i.i.d. draw from the weighted
                // cloud, unrelated to the particle's own parent, and a different one every iteration as the invalid
                // set shrinks.  Same law here: the parent of a uniformly random slot, fresh every round.
                const double u = philox_uniform_elem(p.seed_v, off, slot);
                int64_t sl = static_cast<int64_t>(u * static_cast<double>(p.n_new));
                if (sl >= p.n_new) sl = p.n_new - 1;
                parent = p.parents[sl];
            }
#pragma unroll
            for (int c = 0; c < D; ++c) mu[c] = (p.a * __ldg(p.x_old + parent * D + c)) + msl[c];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double z = 0.0;
#pragma unroll
                for (int m = 0; m < D; ++m)
                    z = fma(Sl[c * D + m], philox_normal_elem(p.seed_n, off, static_cast<int64_t>(m) * p.n_new + slot), z);
                out[c] = mu[c] + z;
            }
            auto row = [&](int c) { return out[c]; };
            ok = model_valid(p.mv, row);
            used = j + 1;
        }
        double* dst = slot_ptr<D>(p, slot);
#pragma unroll
        for (int c = 0; c < D; ++c) dst[c] = out[c];
        if (ok)
            p.list[r] = -1;
        else
            atomicAdd(p.counters + 2, 1ull);
        atomicMax(p.counters + 3, static_cast<unsigned long long>(used));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
        if (is_last) {
            __threadfence();
            const double bad = static_cast<double>(__ldcg(p.counters + 2));
            const double used = static_cast<double>(__ldcg(p.counters + 3));
            const double listed = static_cast<double>(__ldcg(p.counters));
            if (p.mirror != nullptr) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror), "d"(bad), "d"(used), "d"(listed),
                         "d"(p.tag)
                         : "memory");
            p.counters[2] = 0ull;   // ready for the next retry launch
            p.counters[3] = 0ull;
            *p.ticket = 0u;
        }
    }
}

int validate_model(const qb_model* m);

static int bin_grid(int64_t want, int per_sm) {
    int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (cap > BIN_MAX_GRID) cap = BIN_MAX_GRID;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

}  // namespace qb

using namespace qb;

extern "C" size_t qb_lw_binned_workspace_bytes(int64_t n_old, int64_t n_new) {
    if (n_old < 1 || n_new < 1) return 0;
    return bin_layout(n_old, n_new).total;
}

static int launch_sums(const double* d_x, const double* d_w, const double* d_stats, int64_t n_old, int32_t d,
                       double* d_moments_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes, void* stream,
                       bool device_consts, double a, double h, double zero_cov_comp) {
    QB_REQUIRE(d_x && d_w && d_stats && d_ws, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_sums: NULL pointer argument");
    QB_REQUIRE(d >= 1 && d <= 4, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_sums: needs 1 <= d <= 4, got %d", d);
    QB_REQUIRE(n_old >= 1 && n_old < (1LL << 31), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_sums: particle count must lie in [1, 2^31)");
    QB_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_sums: x and w must be 16-byte aligned");
    const BinLayout L = bin_layout(n_old, 1);
    QB_REQUIRE(ws_bytes >= L.total, QB_ERR_WORKSPACE, "qb_lw_binned_sums: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    unsigned int* hdr = reinterpret_cast<unsigned int*>(ws);
    BinSumsParams sp;
    sp.x = d_x;
    sp.w = d_w;
    sp.stats = d_stats;
    sp.n = n_old;
    sp.T = static_cast<int32_t>(L.T);
    sp.pad = 0;
    sp.bounds = reinterpret_cast<double*>(ws + L.bounds);
    sp.counts = reinterpret_cast<unsigned int*>(ws + L.counts);
    sp.partials = reinterpret_cast<double*>(ws + L.partials);
    sp.moments_out = d_moments_out;
    sp.mirror = h_mirror;
    sp.tag = tag;
    sp.ticket = hdr + 0;
    sp.consts = device_consts ? reinterpret_cast<double*>(ws + 128) : nullptr;
    sp.a = a;
    sp.h = h;
    sp.zero_cov_comp = zero_cov_comp;
    static int tma_mode = -1;
    if (tma_mode < 0) {
        const char* e = getenv("QB_BINNED_SUMS_TMA");   // experiment knob: 0 = the register-streaming kernel
        tma_mode = e ? atoi(e) : 1;
    }
    const int64_t full_bins = n_old / BIN;
    if (tma_mode && full_bins >= 1) {
        typedef void (*sums_kernel_t)(const BinSumsParams);
        static const sums_kernel_t kernels[4] = {binned_sums_tma_kernel<1, 3>, binned_sums_tma_kernel<2, 3>,
                                                 binned_sums_tma_kernel<3, 2>, binned_sums_tma_kernel<4, 2>};
        static const int stages[4] = {3, 3, 2, 2};
        static int per_sm[4] = {0, 0, 0, 0};
        const size_t smem = 128 + static_cast<size_t>(stages[d - 1]) * BIN * (d + 1) * 8;
        if (per_sm[d - 1] == 0) {
            QB_CUDA_CHECK(cudaFuncSetAttribute(kernels[d - 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            int occ = 0;
            QB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernels[d - 1], SUMS_THREADS, smem));
            per_sm[d - 1] = occ < 1 ? 1 : occ;
        }
        const int g1 = bin_grid(full_bins, per_sm[d - 1]);
        kernels[d - 1]<<<g1, SUMS_THREADS, smem, st>>>(sp);
        QB_CUDA_CHECK(cudaGetLastError());
        return QB_OK;
    }
    const int g1 = bin_grid((L.T + BIN_THREADS / 32 - 1) / (BIN_THREADS / 32), (d <= 2) ? 4 : 2);
    switch (d) {
        case 1: binned_sums_kernel<1><<<g1, BIN_THREADS, 0, st>>>(sp); break;
        case 2: binned_sums_kernel<2><<<g1, BIN_THREADS, 0, st>>>(sp); break;
        case 3: binned_sums_kernel<3><<<g1, BIN_THREADS, 0, st>>>(sp); break;
        default: binned_sums_kernel<4><<<g1, BIN_THREADS, 0, st>>>(sp); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_binned_sums(const double* d_x, const double* d_w, const double* d_stats, int64_t n_old, int32_t d,
                                 double* d_moments_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                                 void* stream) {
    return launch_sums(d_x, d_w, d_stats, n_old, d, d_moments_out, h_mirror, tag, d_ws, ws_bytes, stream, false, 0.0,
                       0.0, 0.0);
}

// Sharded cloud: global moments = rows summed in rank order, then the same constants the single-GPU pass 1 derives.
template <int D>
__global__ void shard_consts_kernel(const double* __restrict__ rows, int n_ranks, double a, double h,
                                    double zero_cov_comp, double* __restrict__ consts, double* mirror, double tag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int NM = 1 + D + D * D;
    double out[NM];
    for (int k = 0; k < NM; ++k) out[k] = rows[k];
    for (int r = 1; r < n_ranks; ++r)
        for (int k = 0; k < NM; ++k) out[k] += rows[r * NM + k];
    double flags = 0.0, err = 0.0;
    liu_west_consts<D>(out, a, h, zero_cov_comp, consts, flags, err);
    if (mirror != nullptr) {
        for (int k = 0; k < NM; ++k) mirror[k] = out[k];
        mirror[29] = flags;
        mirror[30] = err;
        __threadfence_system();
        *reinterpret_cast<volatile double*>(mirror + 31) = tag;
    }
}

extern "C" int qb_lw_binned_shard_consts(const double* d_rows, int32_t n_ranks, int32_t d, double a, double h,
                                         double zero_cov_comp, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                                         void* stream) {
    QB_REQUIRE(d_rows && d_ws && n_ranks >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_shard_consts: bad arguments");
    QB_REQUIRE(d >= 1 && d <= 4, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_shard_consts: needs 1 <= d <= 4, got %d", d);
    QB_REQUIRE(ws_bytes >= 512, QB_ERR_WORKSPACE, "qb_lw_binned_shard_consts: workspace too small");
    double* consts = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 128);
    cudaStream_t st = as_stream(stream);
    switch (d) {
        case 1: shard_consts_kernel<1><<<1, 32, 0, st>>>(d_rows, n_ranks, a, h, zero_cov_comp, consts, h_mirror, tag); break;
        case 2: shard_consts_kernel<2><<<1, 32, 0, st>>>(d_rows, n_ranks, a, h, zero_cov_comp, consts, h_mirror, tag); break;
        case 3: shard_consts_kernel<3><<<1, 32, 0, st>>>(d_rows, n_ranks, a, h, zero_cov_comp, consts, h_mirror, tag); break;
        default: shard_consts_kernel<4><<<1, 32, 0, st>>>(d_rows, n_ranks, a, h, zero_cov_comp, consts, h_mirror, tag); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_binomial_sample(int64_t n, double p, int64_t count, uint64_t seed, uint64_t off, int64_t* d_out,
                                  void* stream) {
    QB_REQUIRE(d_out && count >= 1 && n >= 0, QB_ERR_INVALID_ARGUMENT, "qb_binomial_sample: bad arguments");
    binomial_test_kernel<<<bin_grid((count + 255) / 256, 8), 256, 0, as_stream(stream)>>>(
        seed, off, static_cast<long long>(n), p, count, reinterpret_cast<long long*>(d_out));
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_binned_count(int64_t n_old, int64_t n_new, uint64_t seed_u, uint64_t off_u, int32_t mode,
                                  void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_ws, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_count: NULL workspace");
    QB_REQUIRE(n_old >= 1 && n_new >= 1 && n_old < (1LL << 31) && n_new < (1LL << 31), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_count: particle counts must lie in [1, 2^31)");
    const BinLayout L = bin_layout(n_old, n_new);
    QB_REQUIRE(ws_bytes >= L.total, QB_ERR_WORKSPACE, "qb_lw_binned_count: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    unsigned int* hdr = reinterpret_cast<unsigned int*>(ws);
    BinCountParams cp;
    cp.bounds = reinterpret_cast<const double*>(ws + L.bounds);
    cp.counts = reinterpret_cast<unsigned int*>(ws + L.counts);
    cp.offs = reinterpret_cast<int64_t*>(ws + L.offs);
    cp.segs = reinterpret_cast<uint2*>(ws + L.segs);
    cp.ticket = hdr + 1;
    cp.nseg = hdr + 3;
    cp.n_new = n_new;
    cp.T = static_cast<int32_t>(L.T);
    cp.max_segs = static_cast<int32_t>(L.max_segs);
    cp.seed_u = seed_u;
    cp.off_u = off_u;
    QB_REQUIRE(mode == QB_COUNT_AUTO || mode == QB_COUNT_TREE || mode == QB_COUNT_HISTOGRAM, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_count: unknown mode %d", mode);
    const size_t smem = static_cast<size_t>(L.T + 2) * 8 + static_cast<size_t>((L.T + 2) & ~1LL) * 4 + 16;
    // measured (B200): while the histogram's tables fit in shared memory it costs ~6 ns per 1000 draws (61 us at
    // n_new = 1e7) and the tree ~10 us per level (143 us for the 13 levels of 1e7 / 2048 bins); beyond that (n > 3.4e7)
    // the histogram falls back to global tables and atomics (1.35 ms at 1e8) and the tree wins
    if (mode == QB_COUNT_AUTO) mode = (smem <= 200 * 1024 && n_new >= 4 * L.T) ? QB_COUNT_HISTOGRAM : QB_COUNT_TREE;
    if (mode == QB_COUNT_TREE) {
        TreeParams tp;
        tp.c = cp;
        tp.heap = reinterpret_cast<unsigned int*>(ws + L.heap);
        tp.P = static_cast<int32_t>(L.P);
        tp.levels = 0;
        while ((1LL << tp.levels) < L.P) ++tp.levels;
        const int top = (tp.levels <= 11) ? tp.levels : 10;
        binned_tree_count_kernel<<<1, COUNT_THREADS, 0, st>>>(tp, top);
        QB_CUDA_CHECK(cudaGetLastError());
        if (top < tp.levels) {
            binned_subtree_kernel<<<1 << top, 64, 0, st>>>(tp, top);
            QB_CUDA_CHECK(cudaGetLastError());
            binned_leaf_scan_kernel<<<1, COUNT_THREADS, 0, st>>>(tp);
            QB_CUDA_CHECK(cudaGetLastError());
        }
        return QB_OK;
    }
    const int64_t npairs = (n_new + 1) / 2;
    if (smem <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            QB_CUDA_CHECK(cudaFuncSetAttribute(binned_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
            attr_set = true;
        }
        const int g2 = bin_grid((npairs + COUNT_THREADS * 4 - 1) / (COUNT_THREADS * 4), 1);
        binned_count_kernel<true><<<g2, COUNT_THREADS, smem, st>>>(cp);
    } else {
        binned_count_kernel<false><<<bin_grid((npairs + COUNT_THREADS - 1) / COUNT_THREADS, 2), COUNT_THREADS, 0, st>>>(cp);
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_binned_prepare(const double* d_x, const double* d_w, const double* d_stats, int64_t n_old,
                                    int32_t d, int64_t n_new, uint64_t seed_u, uint64_t off_u, int32_t count_mode,
                                    double* d_moments_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                                    void* stream) {
    QB_REQUIRE(n_new >= 1 && n_new < (1LL << 31), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_prepare: n_new must lie in [1, 2^31)");
    QB_REQUIRE(n_old >= 1 && ws_bytes >= bin_layout(n_old, n_new).total, QB_ERR_WORKSPACE,
               "qb_lw_binned_prepare: workspace too small");
    int rc = qb_lw_binned_sums(d_x, d_w, d_stats, n_old, d, d_moments_out, h_mirror, tag, d_ws, ws_bytes, stream);
    if (rc != QB_OK) return rc;
    return qb_lw_binned_count(n_old, n_new, seed_u, off_u, count_mode, d_ws, ws_bytes, stream);
}

static int fill_move(BinMoveParams& q, const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                     const double* h_mean, const double* h_S, double a, uint64_t seed_n, uint64_t off_n, int64_t n_new,
                     double* d_x_new, int64_t split, double* d_x_new2, int64_t* d_list,
                     double* h_mirror, double tag, void* d_ws, size_t ws_bytes, const BinLayout& L) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE((h_mean == nullptr) == (h_S == nullptr), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_move/retry: pass both h_mean and h_S, or neither (constants derived by qb_lw_binned_resample)");
    QB_REQUIRE(d_x_old && d_x_new && d_ws, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_move/retry: NULL pointer argument");
    QB_REQUIRE(d == model->d && d >= 1 && d <= 4, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move/retry: needs 1 <= d <= 4");
    QB_REQUIRE(n_old >= 1 && n_new >= 1 && n_old < (1LL << 31) && n_new < (1LL << 31), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_move/retry: particle counts must lie in [1, 2^31)");
    QB_REQUIRE(ws_bytes >= L.total, QB_ERR_WORKSPACE, "qb_lw_binned_move/retry: workspace too small");
    QB_REQUIRE(h_mirror == nullptr || (reinterpret_cast<uintptr_t>(h_mirror) & 31) == 0, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_move/retry: the host mirror must be 32-byte aligned");
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    unsigned int* hdr = reinterpret_cast<unsigned int*>(ws);
    q.x_old = d_x_old;
    q.w = nullptr;
    q.stats = nullptr;
    q.offs = reinterpret_cast<const int64_t*>(ws + L.offs);
    q.segs = reinterpret_cast<const uint2*>(ws + L.segs);
    q.nseg = hdr + 3;
    q.x_new = d_x_new;
    q.x_new2 = d_x_new2;
    q.split = split;
    q.w_new = nullptr;
    q.w_value = 0.0;
    q.stats_new = nullptr;
    q.n_global = 0.0;
    q.list = reinterpret_cast<long long*>(d_list);
    q.counters = reinterpret_cast<unsigned long long*>(ws + 64);
    q.js_out = nullptr;
    q.parents = nullptr;
    q.own_mean = 1;
    q.pad1 = 0;
    q.mirror = h_mirror;
    q.tag = tag;
    q.ticket = hdr + 2;
    q.n_old = n_old;
    q.n_new = n_new;
    q.round_stride = 0;
    q.rounds = 1;
    q.pad0 = 0;
    q.T = static_cast<int32_t>(L.T);
    q.postselect = 1;
    q.seed_v = 0;
    q.off_v = 0;
    q.seed_n = seed_n;
    q.off_n = off_n;
    q.a = a;
    const double oma = 1.0 - a;
    for (int j = 0; j < 16; ++j) q.S[j] = (h_S != nullptr && j < d * d) ? h_S[j] : 0.0;
    for (int c = 0; c < 4; ++c) q.ms[c] = (h_mean != nullptr && c < d) ? oma * h_mean[c] : 0.0;  // (1 - a) * mean
    q.consts = (h_mean == nullptr) ? reinterpret_cast<const double*>(ws + 128) : nullptr;
    q.mv = make_model_view(*model);
    return QB_OK;
}

typedef void (*bin_kernel_t)(const BinMoveParams);

static int launch_retry(const BinMoveParams& q, int d, cudaStream_t st) {
    static const bin_kernel_t kernels[4] = {binned_retry_kernel<1>, binned_retry_kernel<2>, binned_retry_kernel<3>,
                                            binned_retry_kernel<4>};
    const int grid = bin_grid(1 << 20, 1);   // the list length is only known on the device: grid-stride (one CTA per SM)
    kernels[d - 1]<<<grid, 128, 0, st>>>(q);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_lw_binned_move(const qb_model* model, const double* d_x_old, const double* d_w,
                                 const double* d_stats, int64_t n_old, int32_t d, const double* h_mean,
                                 const double* h_S, double a, uint64_t seed_v, uint64_t off_v, uint64_t seed_n,
                                 uint64_t off_n, int64_t n_new, double* d_x_new, int64_t split, double* d_x_new2,
                                 double* d_w_new, int64_t n_global, double* d_stats_new, int32_t postselect,
                                 int32_t retry_rounds, int32_t own_mean, int64_t* d_list, int32_t* d_parents,
                                 int64_t* d_js_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                                 void* stream) {
    BinMoveParams q;
    const BinLayout L = bin_layout(n_old < 1 ? 1 : n_old, n_new < 1 ? 1 : n_new);
    int rc = fill_move(q, model, d_x_old, n_old, d, h_mean, h_S, a, seed_n, off_n, n_new, d_x_new, split, d_x_new2,
                       d_list, h_mirror, tag, d_ws, ws_bytes, L);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_w && d_stats, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move: NULL weights");
    QB_REQUIRE(!postselect || d_list, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move: NULL invalid list");
    QB_REQUIRE(d_w_new == nullptr || n_global >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move: bad n_global");
    QB_REQUIRE(retry_rounds >= 0, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move: negative retry_rounds");
    QB_REQUIRE((reinterpret_cast<uintptr_t>(d_x_old) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_move: x_old and w must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    q.w = d_w;
    q.stats = d_stats;
    q.w_new = d_w_new;
    q.w_value = (n_global >= 1) ? 1.0 / static_cast<double>(n_global) : 0.0;
    q.stats_new = d_stats_new;
    q.n_global = static_cast<double>(n_global);
    q.js_out = d_js_out;
    q.parents = postselect ? d_parents : nullptr;
    q.own_mean = (own_mean || d_parents == nullptr) ? 1 : 0;
    q.postselect = postselect ? 1 : 0;
    q.seed_v = seed_v;
    q.off_v = off_v;
    const size_t smem = static_cast<size_t>(BIN) * (1 + d) * sizeof(double);
    static const bin_kernel_t kernels[4] = {binned_move_kernel<1>, binned_move_kernel<2>, binned_move_kernel<3>,
                                            binned_move_kernel<4>};
    static int per_sm_cache[4] = {0, 0, 0, 0};
    if (per_sm_cache[d - 1] == 0) {
        QB_CUDA_CHECK(cudaFuncSetAttribute(kernels[d - 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        int occ = 0;
        QB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernels[d - 1], BIN_THREADS, smem));
        per_sm_cache[d - 1] = occ < 1 ? 1 : occ;
    }
    const int grid = bin_grid(L.max_segs, per_sm_cache[d - 1]);
    kernels[d - 1]<<<grid, BIN_THREADS, smem, st>>>(q);
    QB_CUDA_CHECK(cudaGetLastError());
    if (postselect && retry_rounds > 0) {
        // the retry launch reads the list length on the device: queue it now, no host round trip in between
        q.rounds = retry_rounds;
        q.round_stride = static_cast<uint64_t>((static_cast<int64_t>(d) * n_new + 1) / 2);
        q.off_n = off_n + q.round_stride;
        q.mirror = (h_mirror != nullptr) ? h_mirror + 4 : nullptr;
        return launch_retry(q, d, st);
    }
    return QB_OK;
}

extern "C" int qb_lw_binned_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                                  const double* h_mean, const double* h_S, double a, uint64_t seed_n, uint64_t off_n,
                                  uint64_t round_stride, int32_t rounds, int32_t own_mean, uint64_t seed_v, int64_t n_new,
                                  double* d_x_new, int64_t split, double* d_x_new2, int64_t* d_list,
                                  const int32_t* d_parents, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                                  void* stream) {
    BinMoveParams q;
    const BinLayout L = bin_layout(n_old < 1 ? 1 : n_old, n_new < 1 ? 1 : n_new);
    int rc = fill_move(q, model, d_x_old, n_old, d, h_mean, h_S, a, seed_n, off_n, n_new, d_x_new, split, d_x_new2,
                       d_list, h_mirror, tag, d_ws, ws_bytes, L);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_list && rounds >= 1, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_retry: bad arguments");
    q.rounds = rounds;
    q.round_stride = round_stride;
    q.parents = const_cast<int32_t*>(d_parents);
    q.own_mean = (own_mean || d_parents == nullptr) ? 1 : 0;
    q.seed_v = seed_v;
    return launch_retry(q, d, as_stream(stream));
}

// The whole device-RNG resample queued by ONE call: pass 1 (sums, moments, Liu-West constants derived on the
// device), pass 2 (counts), pass 3 (move, reading the constants from the workspace) and the first retry launch.
extern "C" int qb_lw_binned_resample(const qb_model* model, const double* d_x, const double* d_w, const double* d_stats,
                                     int64_t n_old, int32_t d, int64_t n_new, double a, double h, double zero_cov_comp,
                                     uint64_t seed, uint64_t off_u, int32_t count_mode, uint64_t off_v, uint64_t seed_n,
                                     uint64_t off_n, double* d_x_new, double* d_w_new, int64_t n_global, double* d_stats_new,
                                     int32_t postselect, int32_t retry_rounds, int32_t own_mean, int64_t* d_list,
                                     int32_t* d_parents, double* d_moments_out, double* h_mirror, double tag,
                                     void* d_ws, size_t ws_bytes, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(n_new >= 1 && n_new < (1LL << 31) && n_old >= 1, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_resample: particle counts must lie in [1, 2^31)");
    QB_REQUIRE(d == model->d, QB_ERR_INVALID_ARGUMENT, "qb_lw_binned_resample: d does not match the model");
    QB_REQUIRE(h_mirror != nullptr && (reinterpret_cast<uintptr_t>(h_mirror) & 31) == 0, QB_ERR_INVALID_ARGUMENT,
               "qb_lw_binned_resample: needs a 32-byte aligned, device-accessible result block of 40 doubles");
    QB_REQUIRE(ws_bytes >= bin_layout(n_old, n_new).total, QB_ERR_WORKSPACE,
               "qb_lw_binned_resample: workspace too small");
    rc = launch_sums(d_x, d_w, d_stats, n_old, d, d_moments_out, h_mirror, tag, d_ws, ws_bytes, stream, true, a, h,
                     zero_cov_comp);
    if (rc != QB_OK) return rc;
    rc = qb_lw_binned_count(n_old, n_new, seed, off_u, count_mode, d_ws, ws_bytes, stream);
    if (rc != QB_OK) return rc;
    return qb_lw_binned_move(model, d_x, d_w, d_stats, n_old, d, nullptr, nullptr, a, seed, off_v, seed_n, off_n, n_new,
                             d_x_new, n_new, nullptr, d_w_new, n_global, d_stats_new, postselect, retry_rounds, own_mean,
                             d_list, d_parents, nullptr, h_mirror + 32, tag, d_ws, ws_bytes, stream);
}

// =====================================================================================================================
// Small clouds (n <= QB_SMALL_MAX), parity mode: the WHOLE first Liu-West pass in ONE single-CTA launch.
// A cloud of 10^3 particles (the reference's own CPU-sized runs, BASELINE config C1) is launch-latency bound: the
// staged parity path costs ~15 launches and 3 host round trips per resample.  Here one CTA holds the weights in shared
// memory and does, in order: moments (block reduction) -> covariance, zero-norm replacement, matrix square root
// (liu_west_consts, as pass 1 of the binned resample) -> np.cumsum of the normalised weights by ONE lane, strictly
// sequential fp64 (resamplers.py:308) -> js = min(searchsorted(cdf, u, 'right'), n - 1) for the HOST-drawn uniforms
// (resamplers.py:318-321; the legacy np.random stream stays on the host, where it is cheapest at this size) ->
// mu = a x[js] + (1 - a) mean, x' = mu + S eps for the host-drawn normals (resamplers.py:325-332, the arithmetic of
// lw_move_small_kernel) -> validity flags and their count -> uniform weights + their stats block.  The retry
// iterations (rare) go through the staged kernels, which find js, the flags and the count where they expect them.
// =====================================================================================================================
constexpr int SMALL_MAX = QB_SMALL_MAX;
constexpr int SMALL_THREADS = 1024;

struct SmallParams {
    const double* x;
    const double* w;
    const double* stats;
    const double* u;        // n_new uniforms
    const double* eps;      // (d, n_new) row-major normals
    double* x_new;
    int64_t* js;
    uint8_t* invalid;
    unsigned long long* counters;   // [0] invalid, [1] clamped draws
    double* w_new;          // may be NULL
    double* stats_new;      // may be NULL
    double* moments_out;    // may be NULL (device copy of the moments)
    double* mirror;         // pinned: [0..) moments, [29] flag, [30] err, [31] tag | [32] invalid, [33] clamped, [34] n_new,
                            // [35] tag | [40..56) S (scaled by h), [56..60) (1 - a) mean
    double tag;
    int32_t n_old, n_new, postselect, pad;
    double a, h, zero_cov_comp;
    ModelView mv;
};

template <int D>
__global__ void __launch_bounds__(SMALL_THREADS) lw_small_resample_kernel(const __grid_constant__ SmallParams p) {
    constexpr int NOUT = 1 + D + D * (D + 1) / 2;
    constexpr int NW = SMALL_THREADS / 32;
    __shared__ double cdf[SMALL_MAX];
    __shared__ double red[NW * NOUT];
    __shared__ double consts[20];
    __shared__ unsigned int s_bad, s_over;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = p.n_old;
    const double inv = p.stats[QB_STAT_INV_NORM];
    if (tid == 0) {
        s_bad = 0u;
        s_over = 0u;
    }
    // ---- normalised weights into shared memory + moment sums ----
    double acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.0;
    for (int i = tid; i < n; i += SMALL_THREADS) {
        const double wv = p.w[i] * inv;
        cdf[i] = wv;
        double xv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xv[c] = p.x[static_cast<size_t>(i) * D + c];
        acc[0] += wv;
        int o = 1 + D;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const double wx = wv * xv[m];
            acc[1 + m] += wx;
#pragma unroll
            for (int c = m; c < D; ++c) acc[o++] += wx * xv[c];
        }
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[wid * NOUT + k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        // finished moments (fixed order) and the Liu-West constants
        double out[1 + D + D * D];
        double fin[NOUT];
        for (int k = 0; k < NOUT; ++k) {
            double v = 0.0;
            for (int q = 0; q < NW; ++q) v += red[q * NOUT + k];
            fin[k] = v;
        }
        out[0] = fin[0];
#pragma unroll
        for (int m = 0; m < D; ++m) out[1 + m] = fin[1 + m];
        int o = 1 + D;
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int c = m; c < D; ++c) {
                out[1 + D + m * D + c] = fin[o];
                out[1 + D + c * D + m] = fin[o];
                ++o;
            }
        double flags = 0.0, err = 0.0;
        liu_west_consts<D>(out, p.a, p.h, p.zero_cov_comp, consts, flags, err);
        if (p.moments_out != nullptr)
            for (int k = 0; k < 1 + D + D * D; ++k) p.moments_out[k] = out[k];
        for (int k = 0; k < 1 + D + D * D; ++k) p.mirror[k] = out[k];
        for (int k = 0; k < 20; ++k) p.mirror[40 + k] = consts[k];
        p.mirror[29] = flags;
        p.mirror[30] = err;
        __threadfence_system();
        *reinterpret_cast<volatile double*>(p.mirror + 31) = p.tag;   // the host starts its checks now
    } else if (tid == 32) {
        // np.cumsum, one lane, in place (the other warps wait at the barrier)
        double run = cdf[0];
        for (int i = 1; i < n; ++i) {
            run = run + cdf[i];
            cdf[i] = run;
        }
    }
    __syncthreads();
    double S[D * D], ms[D];
#pragma unroll
    for (int j = 0; j < D * D; ++j) S[j] = consts[j];
#pragma unroll
    for (int c = 0; c < D; ++c) ms[c] = consts[16 + c];
    // ---- draw + move + validity ----
    const int n_new = p.n_new;
    const int nround = ((n_new + 31) / 32) * 32;
    const double w_value = 1.0 / static_cast<double>(n_new);
    for (int i = tid; i < nround; i += SMALL_THREADS) {
        const bool live = i < n_new;
        bool ok = true, over = false;
        if (live) {
            const double u = p.u[i];
            int lo = 0, hi = n;
            while (lo < hi) {                       // searchsorted(..., side='right')
                const int mid = (lo + hi) >> 1;
                if (u < cdf[mid]) hi = mid; else lo = mid + 1;
            }
            if (lo >= n) {
                lo = n - 1;
                over = true;
            }
            p.js[i] = lo;
            double xv[D], ev[D], out[D];
#pragma unroll
            for (int c = 0; c < D; ++c) xv[c] = p.x[static_cast<size_t>(lo) * D + c];
#pragma unroll
            for (int m = 0; m < D; ++m) ev[m] = p.eps[static_cast<size_t>(m) * n_new + i];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double z = 0.0;
#pragma unroll
                for (int m = 0; m < D; ++m) z = fma(S[c * D + m], ev[m], z);
                out[c] = ((p.a * xv[c]) + ms[c]) + z;   // resamplers.py:325,332, one rounding per ufunc
            }
#pragma unroll
            for (int c = 0; c < D; ++c) p.x_new[static_cast<size_t>(i) * D + c] = out[c];
            if (p.postselect) {
                auto row = [&](int c) { return out[c]; };
                ok = model_valid(p.mv, row);
            }
            p.invalid[i] = ok ? 0 : 1;
            if (p.w_new != nullptr) p.w_new[i] = w_value;
        }
        const unsigned int bad = __ballot_sync(0xffffffffu, live && !ok);
        const unsigned int ovr = __ballot_sync(0xffffffffu, live && over);
        if (lane == 0) {
            if (bad) atomicAdd(&s_bad, static_cast<unsigned int>(__popc(bad)));
            if (ovr) atomicAdd(&s_over, static_cast<unsigned int>(__popc(ovr)));
        }
    }
    __syncthreads();
    if (tid == 0) {
        p.counters[0] = s_bad;
        p.counters[1] = s_over;
        if (p.stats_new != nullptr) {     // resamplers.py:390-392: weights 1/n; the stats block of set_uniform_kernel
            double* st = p.stats_new;
            st[QB_STAT_NORM] = 1.0;
            st[QB_STAT_SUMSQ] = w_value;
            st[QB_STAT_MIN] = w_value;
            st[QB_STAT_NBAD] = 0.0;
            st[QB_STAT_INV_NORM] = 1.0;
            st[QB_STAT_NESS] = static_cast<double>(n_new);
            st[QB_STAT_TAG] = 0.0;
            st[QB_STAT_SKIPPED] = 0.0;
            st[QB_STAT_ATTN] = 0.0;
        }
        p.mirror[32] = static_cast<double>(s_bad);
        p.mirror[33] = static_cast<double>(s_over);
        p.mirror[34] = static_cast<double>(n_new);
        __threadfence_system();
        *reinterpret_cast<volatile double*>(p.mirror + 35) = p.tag;
    }
}

extern "C" int qb_lw_small_resample(const qb_model* model, const double* d_x, const double* d_w, const double* d_stats,
                                    int64_t n_old, int32_t d, double a, double h, double zero_cov_comp,
                                    const double* d_u, const double* d_eps, int64_t n_new, double* d_x_new,
                                    int64_t* d_js, uint8_t* d_invalid, int64_t* d_counters, double* d_w_new,
                                    double* d_stats_new, int32_t postselect, double* d_moments_out, double* h_mirror,
                                    double tag, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(d_x && d_w && d_stats && d_u && d_eps && d_x_new && d_js && d_invalid && d_counters && h_mirror,
               QB_ERR_INVALID_ARGUMENT, "qb_lw_small_resample: NULL pointer argument");
    QB_REQUIRE(d == model->d && d >= 1 && d <= 4, QB_ERR_INVALID_ARGUMENT, "qb_lw_small_resample: needs 1 <= d <= 4");
    QB_REQUIRE(n_old >= 1 && n_old <= SMALL_MAX && n_new >= 1 && n_new < (1LL << 30), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_small_resample: needs 1 <= n_old <= %d (got %lld)", SMALL_MAX, static_cast<long long>(n_old));
    QB_REQUIRE((d_w_new == nullptr) == (d_stats_new == nullptr), QB_ERR_INVALID_ARGUMENT,
               "qb_lw_small_resample: pass both d_w_new and d_stats_new, or neither");
    SmallParams p;
    p.x = d_x;
    p.w = d_w;
    p.stats = d_stats;
    p.u = d_u;
    p.eps = d_eps;
    p.x_new = d_x_new;
    p.js = d_js;
    p.invalid = d_invalid;
    p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    p.w_new = d_w_new;
    p.stats_new = d_stats_new;
    p.moments_out = d_moments_out;
    p.mirror = h_mirror;
    p.tag = tag;
    p.n_old = static_cast<int32_t>(n_old);
    p.n_new = static_cast<int32_t>(n_new);
    p.postselect = postselect ? 1 : 0;
    p.pad = 0;
    p.a = a;
    p.h = h;
    p.zero_cov_comp = zero_cov_comp;
    p.mv = make_model_view(*model);
    cudaStream_t st = as_stream(stream);
    switch (d) {
        case 1: lw_small_resample_kernel<1><<<1, SMALL_THREADS, 0, st>>>(p); break;
        case 2: lw_small_resample_kernel<2><<<1, SMALL_THREADS, 0, st>>>(p); break;
        case 3: lw_small_resample_kernel<3><<<1, SMALL_THREADS, 0, st>>>(p); break;
        default: lw_small_resample_kernel<4><<<1, SMALL_THREADS, 0, st>>>(p); break;
    }
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
