// Fused Bayes-update kernel: likelihood x weight multiply x normalisation sum x
// n_ess reduction in ONE launch (smc.py:324-386,413-453).
//
// HBM traffic per particle-update: read x (8d) + read w (8) + write w' (8)
// = 8(d+2) bytes; the division by the normalisation (smc.py:373) is deferred
// as the scalar stats[INV_NORM] that the NEXT pass folds into its weight load.
//
// Data movement: full tiles of the row-major (n, d) particle slab and of the
// weight vector are staged global->shared with 1-D bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a STAGES-deep
// ring, so every SM keeps STAGES x 16-24 KB of loads in flight without holding
// them in registers.  For d in {1,3,4} each thread then owns PAIRS of adjacent
// particles: 128-bit conflict-free shared loads in, one 128-bit streaming store
// of the two new weights out, all tile offsets compile-time immediates — the
// instruction count per particle is what bounds this kernel once the memory
// system is fed (ncu r1: 114 -> ~35 warp-instructions per 32 particles).
// Generic d (tomography) walks its row with a per-lane rotated start so that
// rows 128 B apart do not collide on shared-memory banks.
// The ragged last tile (byte count not a multiple of 16) is loaded directly.
#include <cstdlib>
#include "qb_models.cuh"

namespace qb {

constexpr int UPD_CONSUMER_WARPS = 8;
constexpr int UPD_THREADS = (UPD_CONSUMER_WARPS + 1) * 32;  // + one TMA producer warp
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
    double* mirror;          // device-accessible pinned host copy of the stats block (or NULL)
    double tag;
    double zero_weight_thresh, resample_below;
    int32_t guard, guard_resample;
    int32_t n_ranks, rank;   // > 1: all-reduce the three sums over the peers' mailboxes inside this launch
    double* peer_mbox[QB_MAX_RANKS];
    int32_t* error_flag;     // device int set to 1 if the peer wait timed out
    ModelView mv;
    ExpView ev;
    double meas[QB_MAX_D];
};

__device__ __forceinline__ void block_reduce3(double& s, double& q, unsigned int& bad, double* red,
                                              unsigned int* redu) {
    s = warp_sum(s);
    q = warp_sum(q);
    bad = __reduce_add_sync(0xffffffffu, bad);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[wid * 2 + 0] = s;
        red[wid * 2 + 1] = q;
        redu[wid] = bad;
    }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        s = (lane < nw) ? red[lane * 2 + 0] : 0.0;
        q = (lane < nw) ? red[lane * 2 + 1] : 0.0;
        bad = (lane < nw) ? redu[lane] : 0u;
        s = warp_sum(s);
        q = warp_sum(q);
        bad = __reduce_add_sync(0xffffffffu, bad);
    }
}

// Does the step that produced `st` need the host before another update may run?  (negative/NaN
// weights, smc.py:416; the zero-weight policies, smc.py:423-436; the resample trigger, smc.py:275;
// or it was itself skipped.)  Evaluated on the device so that the NEXT update can be launched
// speculatively and cancel itself.
__device__ __forceinline__ bool needs_host(const double* st, double zero_thresh, int check_resample,
                                           double resample_below) {
    const double eps = 2.220446049250313e-16;
    const double norm = st[QB_STAT_NORM];
    const double total = (fabs(norm) < eps) ? norm : 1.0;  // np.sum of the normalised weights
    bool attn = (st[QB_STAT_NBAD] > 0.0) || (total <= zero_thresh) || (st[QB_STAT_SKIPPED] != 0.0);
    if (check_resample) attn = attn || (st[QB_STAT_NESS] < resample_below);
    return attn;
}

__device__ __forceinline__ void publish_stats(const UpdateParams& p, double norm, double sumsq, double nbad,
                                              double skipped) {
    const double eps = 2.220446049250313e-16;  // np.spacing(1), smc.py:370
    double v[QB_STAT_COUNT];
#pragma unroll
    for (int k = 0; k < QB_STAT_COUNT; ++k) v[k] = 0.0;
    v[QB_STAT_NORM] = norm;
    v[QB_STAT_SUMSQ] = sumsq;
    v[QB_STAT_MIN] = nan("");  // computed on demand (qb_weights_min) when NBAD > 0
    v[QB_STAT_NBAD] = nbad;
    v[QB_STAT_INV_NORM] = (fabs(norm) < eps) ? 1.0 : 1.0 / norm;
    // n_ess = 1 / sum(w_normalised^2); when the norm guard of smc.py:369-370 applies the weights stay as they are
    v[QB_STAT_NESS] = (fabs(norm) < eps) ? 1.0 / sumsq : (norm * norm) / sumsq;
    v[QB_STAT_TAG] = p.tag;
    v[QB_STAT_SKIPPED] = skipped;
#pragma unroll
    for (int k = 0; k < 8; ++k) p.stats_out[k] = v[k];
    if (p.mirror != nullptr) {
        // Host mirror without a system-scope fence (a PCIe round trip on the kernel's critical path): the block is
        // written as two 32-byte vector stores, each a single aligned PCIe write, and EACH half carries the tag —
        // slots [3] and [6] — so the host accepts a snapshot only when both tags match (it re-reads otherwise).
        // Layout of the mirror: {NORM, SUMSQ, NBAD, TAG | INV_NORM, NESS, TAG, SKIPPED}.
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror), "d"(v[QB_STAT_NORM]),
                     "d"(v[QB_STAT_SUMSQ]), "d"(v[QB_STAT_NBAD]), "d"(p.tag)
                     : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 4), "d"(v[QB_STAT_INV_NORM]),
                     "d"(v[QB_STAT_NESS]), "d"(p.tag), "d"(skipped)
                     : "memory");
    }
}

// In-kernel all-reduce of (norm, sumsq, nbad) across the ranks of one NVLink domain.  Every rank owns a
// mailbox of 2 x n_ranks x 4 doubles mapped into all peers (CUDA IPC).  Launch `tag` uses half (tag & 1):
// lane q stores this rank's three sums into peer q's mailbox row [half][rank] and then, after a
// system-scope fence, the tag word; it then spins on its own mailbox row [half][q] until peer q's tag
// arrives.  The rows are summed in rank order, so every rank obtains bit-identical global sums.
// Two halves suffice: a rank can only be one launch ahead of the slowest peer (it needs that peer's
// previous-launch row to finish its own previous launch).
__device__ void peer_allreduce3(const UpdateParams& p, double& s, double& q, double& bad, double* red) {
    const int G = p.n_ranks;
    const int half = static_cast<int>(static_cast<long long>(p.tag) & 1LL);
    const int lane = threadIdx.x;
    if (lane < G) {
        volatile double* dst = p.peer_mbox[lane] + (static_cast<size_t>(half) * G + p.rank) * 4;
        dst[0] = s;
        dst[1] = q;
        dst[2] = bad;
        __threadfence_system();
        dst[3] = p.tag;
        volatile double* src = p.peer_mbox[p.rank] + (static_cast<size_t>(half) * G + lane) * 4;
        const long long t0 = clock64();
        bool ok = true;
        while (src[3] != p.tag) {
            if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; do not hang the GPU
                ok = false;
                break;
            }
        }
        __threadfence_system();
        red[lane * 3 + 0] = ok ? src[0] : nan("");
        red[lane * 3 + 1] = ok ? src[1] : nan("");
        red[lane * 3 + 2] = ok ? src[2] : 0.0;
        if (!ok && p.error_flag != nullptr) *p.error_flag = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0, tb = 0.0;
        for (int r = 0; r < G; ++r) {
            ts += red[r * 3 + 0];
            tq += red[r * 3 + 1];
            tb += red[r * 3 + 2];
        }
        s = ts;
        q = tq;
        bad = tb;
    }
}

// Final deterministic reduction of per-block partials by the last block to finish.
__device__ void finish_stats(const UpdateParams& p, int nblocks, double* red, unsigned int* redu, double* redp) {
    double s = 0.0, q = 0.0, bad = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        s += p.partials[b * 4 + 0];
        q += p.partials[b * 4 + 1];
        bad += p.partials[b * 4 + 2];
    }
    unsigned int ubad = static_cast<unsigned int>(bad);
    __syncthreads();
    block_reduce3(s, q, ubad, red, redu);
    double dbad = static_cast<double>(ubad);
    if (p.n_ranks > 1) {
        // broadcast the block totals held by thread 0 to the lanes that talk to the peers
        if (threadIdx.x == 0) {
            redp[3 * QB_MAX_RANKS + 0] = s;
            redp[3 * QB_MAX_RANKS + 1] = q;
            redp[3 * QB_MAX_RANKS + 2] = dbad;
        }
        __syncthreads();
        s = redp[3 * QB_MAX_RANKS + 0];
        q = redp[3 * QB_MAX_RANKS + 1];
        dbad = redp[3 * QB_MAX_RANKS + 2];
        __syncthreads();
        peer_allreduce3(p, s, q, dbad, redp);
    }
    if (threadIdx.x == 0) publish_stats(p, s, q, dbad, 0.0);
}

struct Acc {
    double s, q;
    unsigned int bad;
};

__device__ __forceinline__ void accumulate(Acc& a, double wv) {
    a.s += wv;
    a.q = fma(wv, wv, a.q);
    a.bad += (wv >= 0.0) ? 0u : 1u;  // counts negatives and NaNs (smc.py:416)
}

// DT > 0: compile-time n_modelparams, pair processing.  DT == 0: runtime d.
//
// Warp-specialised: warps 0..UPD_CONSUMER_WARPS-1 compute, the last warp's lane 0 is the TMA producer.
// full[s]  (count 1)  : producer's expect_tx + the bulk copies' complete_tx  -> consumers may read stage s
// empty[s] (count NCW): one arrive per consumer warp                        -> producer may refill stage s
// No CTA-wide barrier in the tile loop: warps drift freely across the ring.
template <int KIND, bool BINOM, int DT>
__global__ void __launch_bounds__(UPD_THREADS, 4) fused_update_kernel(const __grid_constant__ UpdateParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TILE_CT = (DT == 1) ? 1024 : 512;  // must match choose_tile()
    constexpr int NCT = UPD_CONSUMER_WARPS * 32;     // consumer threads
    const int d = (DT > 0) ? DT : p.d;
    const int tile = (DT > 0) ? TILE_CT : p.tile;
    const uint32_t x_bytes = static_cast<uint32_t>(tile) * d * 8u;
    const uint32_t w_bytes = static_cast<uint32_t>(tile) * 8u;
    const uint32_t stage_bytes = x_bytes + w_bytes;  // multiples of 128 by construction (tile % 16 == 0)
    double* meas_s = reinterpret_cast<double*>(smem_raw + 128);
    unsigned char* ring = smem_raw + 128 + QB_MAX_D * 8;
    __shared__ double red[(UPD_THREADS / 32) * 2];
    __shared__ unsigned int redu[UPD_THREADS / 32];
    __shared__ double redp[3 * QB_MAX_RANKS + 3];
    __shared__ unsigned int is_last;

    const int tid = threadIdx.x;
    // Programmatic dependent launch: let the NEXT launch on this stream become resident while this one runs
    // (its CTAs take the slots ours free and park in griddepcontrol.wait), and do not touch anything the
    // PREVIOUS launch wrote (stats_in, w_in, the ticket) before that launch has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.guard && needs_host(p.stats_in, p.zero_weight_thresh, p.guard_resample, p.resample_below)) {
        // speculative launch whose predecessor needs the host: do nothing, say so
        if (blockIdx.x == 0 && tid == 0) publish_stats(p, p.stats_in[QB_STAT_NORM], p.stats_in[QB_STAT_SUMSQ], 0.0, 1.0);
        return;
    }
    const uint32_t bar0 = smem_u32(smem_raw);            // full[s] at bar0 + 8 s, empty[s] at bar0 + 64 + 8 s
    const uint32_t ring0 = smem_u32(ring);
    const int ntiles = static_cast<int>((p.n + tile - 1) / tile);
    const int my_tiles = (ntiles > static_cast<int>(blockIdx.x))
                             ? (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 0;

    if (KIND == QB_MODEL_TOMOGRAPHY) {
        for (int c = tid; c < d; c += UPD_THREADS) meas_s[c] = p.meas[c];
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < UPD_STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 64 + 8 * s, UPD_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    Acc a0 = {0.0, 0.0, 0u}, a1 = {0.0, 0.0, 0u};
    const int64_t tile_stride = static_cast<int64_t>(gridDim.x) * tile;

    if (tid >= NCT) {
        // ===== producer warp: one lane streams my tiles into the ring =====
        if (tid == NCT) {
            int s = 0;
            uint32_t phase = 0;
            int64_t first = static_cast<int64_t>(blockIdx.x) * tile;
            for (int i = 0; i < my_tiles; ++i, first += tile_stride) {
                if (first + tile <= p.n) {  // full tile (the ragged last one is read directly by the consumers)
                    mbar_wait(bar0 + 64 + 8 * s, phase ^ 1u);  // passes at once the first time round the ring
                    const uint32_t dst = ring0 + static_cast<uint32_t>(s) * stage_bytes;
                    mbar_expect_tx(bar0 + 8 * s, stage_bytes);
                    tma_load_1d(dst, p.x + first * d, x_bytes, bar0 + 8 * s);
                    tma_load_1d(dst + x_bytes, p.w_in + first, w_bytes, bar0 + 8 * s);
                }
                if (++s == UPD_STAGES) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ===== consumer warps =====
        const double inv_norm = p.stats_in[QB_STAT_INV_NORM];
        const ModelView mv = p.mv;
        const ExpView ev = p.ev;
        const int lane = tid & 31;
        auto meas = [&](int c) { return meas_s[c]; };
        int s = 0;
        uint32_t phase = 0;
        int64_t first = static_cast<int64_t>(blockIdx.x) * tile;
        for (int i = 0; i < my_tiles; ++i, first += tile_stride) {
            double* wo = p.w_out + first;
            if (first + tile <= p.n) {
                const unsigned char* stage = ring + static_cast<size_t>(s) * stage_bytes;
                const double* xs = reinterpret_cast<const double*>(stage);
                const double* ws = reinterpret_cast<const double*>(stage + x_bytes);
                mbar_wait(bar0 + 8 * s, phase);
                if constexpr (DT > 0) {
                    // pairs (2j, 2j+1): 128-bit shared loads, one 128-bit streaming store
                    constexpr int NPAIRS = TILE_CT >> 1;
#pragma unroll
                    for (int j0 = 0; j0 < NPAIRS; j0 += NCT) {
                        const int j = j0 + tid;
                        const double2 wp = *reinterpret_cast<const double2*>(ws + 2 * j);
                        double xr[2 * DT];
#pragma unroll
                        for (int v = 0; v < DT; ++v) {
                            const double2 t2 = *reinterpret_cast<const double2*>(xs + 2 * DT * j + 2 * v);
                            xr[2 * v] = t2.x;
                            xr[2 * v + 1] = t2.y;
                        }
                        auto row0 = [&](int c) { return xr[c]; };
                        auto row1 = [&](int c) { return xr[DT + c]; };
                        const double L0 = model_likelihood<KIND, BINOM>(mv, ev, row0, meas, 0);
                        const double L1 = model_likelihood<KIND, BINOM>(mv, ev, row1, meas, 0);
                        const double w0 = (wp.x * inv_norm) * L0;  // smc.py:354 on the lazily normalised weight
                        const double w1 = (wp.y * inv_norm) * L1;
                        asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(wo + 2 * j), "d"(w0),
                                     "d"(w1)
                                     : "memory");
                        accumulate(a0, w0);
                        accumulate(a1, w1);
                    }
                } else {
                    for (int j = tid; j < tile; j += NCT) {
                        const double* xr = xs + static_cast<size_t>(j) * d;
                        auto row = [&](int c) { return xr[c]; };
                        const double L = model_likelihood<KIND, BINOM>(mv, ev, row, meas, lane);
                        const double wv = (ws[j] * inv_norm) * L;
                        stg_stream(wo + j, wv);
                        accumulate(a0, wv);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 64 + 8 * s);  // this warp is done with stage s
            } else {  // ragged last tile: straight from global memory
                const int cnt = static_cast<int>(p.n - first);
                for (int j = tid; j < cnt; j += NCT) {
                    const double* xr = p.x + (first + j) * d;
                    auto row = [&](int c) { return xr[c]; };
                    const double L = model_likelihood<KIND, BINOM>(mv, ev, row, meas, 0);
                    const double wv = (p.w_in[first + j] * inv_norm) * L;
                    wo[j] = wv;
                    accumulate(a0, wv);
                }
            }
            if (++s == UPD_STAGES) {
                s = 0;
                phase ^= 1u;
            }
        }
    }

    double acc_s = a0.s + a1.s, acc_q = a0.q + a1.q;
    unsigned int acc_bad = a0.bad + a1.bad;
    block_reduce3(acc_s, acc_q, acc_bad, red, redu);
    if (tid == 0) {
        p.partials[blockIdx.x * 4 + 0] = acc_s;
        p.partials[blockIdx.x * 4 + 1] = acc_q;
        p.partials[blockIdx.x * 4 + 2] = static_cast<double>(acc_bad);
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        finish_stats(p, gridDim.x, red, redu, redp);
        if (tid == 0) *p.ticket = 0u;  // ready for the next launch on this stream
    }
}

// Tile size: 16-24 KB per stage; a multiple of 2 * UPD_THREADS particles for the pair kernels,
// of 16 particles otherwise, so every bulk copy is 128-B granular.
static int choose_tile(int d) {
    if (d == 1) return 1024;   // 16 KB / stage
    if (d == 3 || d == 4) return 512;  // 16 / 20 KB
    int t = 2048 / (d + 1);
    t = (t / 16) * 16;
    if (t < 16) t = 16;
    return t;
}

static size_t update_smem_bytes(int d) {
    return 128 + QB_MAX_D * 8 + static_cast<size_t>(UPD_STAGES) * choose_tile(d) * (d + 1) * 8;
}

typedef void (*update_kernel_t)(const UpdateParams);

static update_kernel_t pick_update_kernel(const qb_model& m) {
    switch (m.kind) {
        case QB_MODEL_PRECESSION:
            return m.binomial ? fused_update_kernel<QB_MODEL_PRECESSION, true, 1>
                              : fused_update_kernel<QB_MODEL_PRECESSION, false, 1>;
        case QB_MODEL_RB:
            if (m.interleaved)
                return m.binomial ? fused_update_kernel<QB_MODEL_RB, true, 4> : fused_update_kernel<QB_MODEL_RB, false, 4>;
            return m.binomial ? fused_update_kernel<QB_MODEL_RB, true, 3> : fused_update_kernel<QB_MODEL_RB, false, 3>;
        case QB_MODEL_TOMOGRAPHY:
            return m.binomial ? fused_update_kernel<QB_MODEL_TOMOGRAPHY, true, 0>
                              : fused_update_kernel<QB_MODEL_TOMOGRAPHY, false, 0>;
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

static int update_grid_limit(update_kernel_t k, size_t smem) {
    // the attribute call is idempotent and cheap; keeping it unconditional avoids per-device caches
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, UPD_THREADS, smem) != cudaSuccess) return -1;
    if (per_sm < 1) per_sm = 1;
    return per_sm * sm_count();
}

struct GridCacheEntry {
    update_kernel_t k;
    int dev;
    int limit;
};
static GridCacheEntry g_grid_cache[32];
static int g_grid_cache_n = 0;

static int cached_grid_limit(update_kernel_t k, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < g_grid_cache_n; ++i)
        if (g_grid_cache[i].k == k && g_grid_cache[i].dev == dev) return g_grid_cache[i].limit;
    const int limit = update_grid_limit(k, smem);
    if (limit > 0 && g_grid_cache_n < 32) g_grid_cache[g_grid_cache_n++] = {k, dev, limit};
    return limit;
}

// ---- smallest weight (only needed for the warning text of smc.py:417) ------------------------
__global__ void __launch_bounds__(256) weights_min_kernel(const double* __restrict__ w, int64_t n, double* out) {
    __shared__ double red[8];
    double mn = INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += 256) mn = fmin(mn, w[i]);
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) mn = fmin(mn, red[k]);
        *out = mn;
    }
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
                               double* d_stats_out, const qb_update_ctl* ctl, void* d_ws, size_t ws_bytes,
                               void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(ep && d_x && d_w_in && d_w_out && d_stats_in && d_stats_out && d_ws, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: NULL pointer argument");
    QB_REQUIRE(n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: n must be >= 1, got %lld", (long long)n);
    QB_REQUIRE(ws_bytes >= qb_update_workspace_bytes(n, model->d), QB_ERR_WORKSPACE,
               "qb_fused_update: workspace too small");
    QB_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w_in) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d_w_out) & 15) == 0,
               QB_ERR_INVALID_ARGUMENT, "qb_fused_update: x, w_in and w_out must be 16-byte aligned");

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
    p.mirror = ctl ? ctl->h_mirror : nullptr;
    p.tag = ctl ? ctl->tag : 0.0;
    p.zero_weight_thresh = ctl ? ctl->zero_weight_thresh : 0.0;
    p.resample_below = ctl ? ctl->resample_below : 0.0;
    p.guard = ctl ? ctl->guard : 0;
    p.guard_resample = ctl ? ctl->guard_resample : 0;
    p.n_ranks = ctl ? ctl->n_ranks : 0;
    p.rank = ctl ? ctl->rank : 0;
    p.error_flag = ctl ? ctl->d_error_flag : nullptr;
    QB_REQUIRE(p.n_ranks <= QB_MAX_RANKS, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: at most %d ranks", QB_MAX_RANKS);
    for (int r = 0; r < QB_MAX_RANKS; ++r) p.peer_mbox[r] = (ctl && r < p.n_ranks) ? ctl->d_peer_mailbox[r] : nullptr;
    if (p.n_ranks > 1) {
        QB_REQUIRE(p.rank >= 0 && p.rank < p.n_ranks, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: bad rank");
        for (int r = 0; r < p.n_ranks; ++r)
            QB_REQUIRE(p.peer_mbox[r] != nullptr, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: NULL peer mailbox %d", r);
    }
    QB_REQUIRE(d_stats_in != d_stats_out || !p.guard, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: a guarded update needs distinct stats_in / stats_out blocks");
    p.mv = make_model_view(*model);
    p.ev = make_exp_view(*model, *ep, outcome);
    for (int c = 0; c < QB_MAX_D; ++c) p.meas[c] = (c < model->d) ? ep->meas[c] : 0.0;

    update_kernel_t k = pick_update_kernel(*model);
    const size_t smem = update_smem_bytes(model->d);
    const int limit = cached_grid_limit(k, smem);
    QB_REQUIRE(limit > 0, QB_ERR_CUDA, "qb_fused_update: occupancy query failed: %s",
               cudaGetErrorString(cudaGetLastError()));
    const int64_t ntiles = (n + p.tile - 1) / p.tile;
    int grid = static_cast<int>(ntiles < limit ? ntiles : limit);
    {
        static int env_cap = -1;  // experiment knob: QB_UPD_CTAS_PER_SM caps the resident CTAs per SM
        if (env_cap < 0) {
            const char* e = getenv("QB_UPD_CTAS_PER_SM");
            env_cap = e ? atoi(e) : 0;
        }
        if (env_cap > 0 && grid > env_cap * sm_count()) grid = env_cap * sm_count();
    }
    if (grid > 32 * 256) grid = 32 * 256;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(UPD_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k, p));
    return QB_OK;
}

extern "C" int qb_weights_min(const double* d_w, int64_t n, double* d_out, void* stream) {
    QB_REQUIRE(d_w && d_out && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_weights_min: bad arguments");
    weights_min_kernel<<<1, 256, 0, as_stream(stream)>>>(d_w, n, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
