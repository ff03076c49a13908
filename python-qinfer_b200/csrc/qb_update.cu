// Fused Bayes-update kernel: likelihood x weight multiply x normalisation sum x
// n_ess reduction in ONE launch (smc.py:324-386,413-453), for ONE update or for a
// batch of up to QB_MAX_FUSE consecutive updates (smc.py:459-487 batch_update).
//
// HBM traffic per launch: read x (8d) + read w (8) + write w' (8) = 8(d+2) bytes per particle,
// whatever the number K of fused updates — the division by the normalisation (smc.py:373) is deferred
// as the scalar stats[INV_NORM] that the NEXT launch folds into its weight load, and the K likelihoods
// of a batch are evaluated on the particle while it sits in registers.  Per-step sums S_j = sum w_j,
// Q_j = sum w_j^2 and bad-weight counts are reduced for every step j of the batch, so the host recovers
// each step's normalisation record (S_j / S_{j-1}), n_ess (S_j^2 / Q_j) and policy tests exactly as if
// the updates had been issued one by one; the weights differ from the one-by-one path only by the
// omitted intermediate renormalisation (rounding, ~K ulp).
//
// Data movement: full tiles of the row-major (n, d) particle slab and of the weight vector are staged
// global->shared with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a
// STAGES-deep full/empty mbarrier ring by one producer lane; 8 consumer warps read them back.  For d in
// {1,3,4} each thread owns PAIRS of adjacent particles: 128-bit conflict-free shared loads in, one 128-bit
// streaming store of the two new weights out.  Generic d (tomography) walks its row with a per-lane rotated
// start so that rows 128 B apart do not collide on shared-memory banks.  The ragged last tile (byte count
// not a multiple of 16) is read directly.  Launched with programmatic stream serialisation: the next
// launch's CTAs become resident (2 CTAs/SM per launch leave room for it) and park in griddepcontrol.wait.
#include <cstdlib>
#include "qb_models.cuh"

namespace qb {

constexpr int UPD_CONSUMER_WARPS = 8;
constexpr int UPD_THREADS = (UPD_CONSUMER_WARPS + 1) * 32;  // + one TMA producer warp
constexpr int UPD_STAGES = 3;
constexpr int KF_MAX = QB_MAX_FUSE;
constexpr int MBOX_ROW = QB_MAILBOX_ROW;  // 8-byte slots per mailbox row (two flagged words per sum)

struct UpdateParams {
    const double* x;
    const double* w_in;
    double* w_out;
    const double* stats_in;
    double* stats_out;
    double* step_stats;      // [nsteps][8] device copy of the per-step blocks (may be NULL)
    double* partials;        // [grid][3 * KF]
    unsigned int* ticket;    // last-block-done counter (self-resetting)
    int64_t n;
    int32_t tile;            // particles per tile
    int32_t d;
    int32_t nsteps;          // 1..KF_MAX updates fused in this launch
    uint32_t resample_mask;  // bit j: step j is followed by an n_ess check (check_for_resample)
    double* mirror;          // device-accessible pinned host block, nsteps x 8 doubles (or NULL)
    double tag;
    double zero_weight_thresh, resample_below;
    int32_t guard;
    int32_t reverse;         // 1: walk the tiles from the END of the slab (zig-zag across launches, see the kernel)
    int32_t l2hint;          // bit 0: w_in evict_first, bit 1: x evict_last, bit 2: w_out evict_last, bit 3: partition
    int32_t pin_tiles;       // partition mode: tiles [0, pin_tiles) of x / w_in / w_out are kept in L2 (evict_last),
                             // the rest streams through (evict_first)
    double chain_prev_tag;   // != 0: the launch right before this one on the stream is the update with that tag on the
                             // same buffers; depend on it through its flags (below) instead of griddepcontrol.wait
    double* data_tag;        // device word: tag of the last launch whose weights are completely written
    int32_t chain_capable, pad2;  // the host chains launches on this cloud: publish the flags with release ordering
    int32_t n_ranks, rank;   // > 1: all-reduce the sums over the peers' mailboxes inside this launch
    double* peer_mbox[QB_MAX_RANKS];
    int32_t* error_flag;     // device int set to 1 if the peer wait timed out
    double* shard_norms;     // pinned host block (or NULL): every rank's own sum w' of the last step + the tag
    ModelView mv;
    ExpView ev[KF_MAX];
    double meas[QB_MAX_D];   // tomography (single-step launches only)
};

struct Acc {
    double s, q;
};

// bit k of `bad` is set if step k produced a negative or NaN weight (smc.py:416 only asks "not all >= 0")
__device__ __forceinline__ void accumulate(Acc& a, unsigned int& bad, int k, double wv) {
    a.s += wv;
    a.q = fma(wv, wv, a.q);
    bad |= (wv >= 0.0) ? 0u : (1u << k);
}

// Block reduction of KF (s, q, bad) triples into out[3*j..3*j+2] (shared memory).
template <int KF>
__device__ __forceinline__ void block_reduce_steps(const Acc (&a)[KF], unsigned int bad, double* red, double* out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const unsigned int wbad = __reduce_or_sync(0xffffffffu, bad);
#pragma unroll
    for (int j = 0; j < KF; ++j) {
        const double s = warp_sum(a[j].s);
        const double q = warp_sum(a[j].q);
        if (lane == 0) {
            red[(wid * KF + j) * 3 + 0] = s;
            red[(wid * KF + j) * 3 + 1] = q;
            red[(wid * KF + j) * 3 + 2] = ((wbad >> j) & 1u) ? 1.0 : 0.0;  // summed: > 0 iff any warp saw one
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * KF) {
        double v = 0.0;
        for (int w = 0; w < nw; ++w) v += red[w * KF * 3 + threadIdx.x];
        out[threadIdx.x] = v;
    }
    __syncthreads();
}

// Does the launch that produced `st` need the host before another update may run?  (negative/NaN
// weights, smc.py:416; the zero-weight policies, smc.py:423-436; the resample trigger, smc.py:275;
// or it cancelled itself.)  The producing launch evaluates those tests itself (publish) so that the NEXT
// launch can be issued speculatively and cancel itself on one flag.
// (L2 loads: with a chained launch the block may have been rewritten while this SM still caches the old lines)
__device__ __forceinline__ bool needs_host(const double* st) {
    return (__ldcg(st + QB_STAT_ATTN) != 0.0) || (__ldcg(st + QB_STAT_SKIPPED) != 0.0);
}

// Publish the per-step blocks and the final stats block.  sums[3*j..] = (S_j, Q_j, nbad_j), global.
__device__ void publish(const UpdateParams& p, const double* sums) {
    const double skipped = 0.0;
    const double eps = 2.220446049250313e-16;  // np.spacing(1), smc.py:370
    const int K = p.nsteps;
    // pass 0 derives the final stats block and releases it (a chained successor is waiting for exactly that);
    // pass 1 writes the per-step blocks (device copy, host mirror) — the host can take another microsecond
    for (int pass = 0; pass < 2; ++pass) {
        double attn = 0.0;
        double s_prev = 1.0;
        double norm = 0.0, sumsq = 0.0, nbad_tot = 0.0, ness = 0.0;
        for (int j = 0; j < K; ++j) {
            const double S = sums[3 * j + 0], Q = sums[3 * j + 1], nb = sums[3 * j + 2];
            // normalisation record of step j (smc.py:357): sum of (normalised previous weights) * L_j
            const double rec = (j == 0) ? S : S / s_prev;
            const bool degenerate = fabs(rec) < eps;      // smc.py:369-370: then the weights stay as w * L
            const double total = degenerate ? rec : 1.0;  // np.sum of the weights the reference would hold
            const double ne = degenerate ? 1.0 / Q : (S * S) / Q;
            bool a = (nb > 0.0) || (total <= p.zero_weight_thresh);
            if ((p.resample_mask >> j) & 1u) a = a || (ne < p.resample_below);
            const double flag = (a ? 1.0 : 0.0) + 2.0 * skipped;
            if (a && attn == 0.0) attn = static_cast<double>(j + 1);
            if (pass == 1 && p.step_stats != nullptr) {
                double* o = p.step_stats + 8 * j;
                o[0] = S;
                o[1] = Q;
                o[2] = nb;
                o[3] = p.tag;
                o[4] = rec;
                o[5] = ne;
                o[6] = p.tag;
                o[7] = flag;
            }
            if (pass == 1 && p.mirror != nullptr) {
                // Host mirror without a system-scope fence: two 32-byte vector stores per step, each a single
                // aligned PCIe write and EACH carrying the tag; the host accepts a block only when both tags match.
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j), "d"(S), "d"(Q),
                             "d"(nb), "d"(p.tag)
                             : "memory");
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j + 4), "d"(rec),
                             "d"(ne), "d"(p.tag), "d"(flag)
                             : "memory");
            }
            s_prev = S;
            norm = S;
            sumsq = Q;
            nbad_tot += nb;
            ness = ne;
        }
        if (pass == 0) {
            double* so = p.stats_out;
            so[QB_STAT_NORM] = norm;    // sum of the stored (unnormalised) weights after the last step
            so[QB_STAT_SUMSQ] = sumsq;
            so[QB_STAT_MIN] = nan("");  // computed on demand (qb_weights_min) when NBAD > 0
            so[QB_STAT_NBAD] = nbad_tot;
            so[QB_STAT_INV_NORM] = (fabs(norm) < eps) ? 1.0 : 1.0 / norm;
            so[QB_STAT_NESS] = ness;
            so[QB_STAT_SKIPPED] = skipped;
            so[QB_STAT_ATTN] = attn;
            // the tag goes last; behind a fence when a successor may be polling it (chained launches)
            if (p.chain_capable) __threadfence();
            *reinterpret_cast<volatile double*>(so + QB_STAT_TAG) = p.tag;
        }
    }
}

// In-kernel all-reduce of the 3K sums across the ranks of one NVLink domain.  Every rank owns a mailbox of
// 2 x n_ranks rows of MBOX_ROW 8-byte slots mapped into all peers (CUDA IPC).  Launch `tag` uses half (tag & 1).
// Flagged-word protocol (the "LL" idea): every double travels as two 8-byte words {32 data bits | 32-bit flag},
// flag = the low 32 bits of the launch tag.  An 8-byte store is atomic, so a word whose flag matches carries
// valid data: no fence between data and flag on the sender, no fence after the flag on the receiver, and the
// data and its validity arrive in ONE NVLink transaction.  Lane (q, k) stores word k of this rank's sums into
// peer q's row [half][rank] and then polls word k of its own row [half][q].  Rows are summed in rank order, so
// every rank obtains bit-identical global sums.  Two halves suffice: a rank can be at most one launch ahead of the
// slowest peer (it needs that peer's previous-launch row to finish its own previous launch).
__device__ __forceinline__ void st_word_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ void peer_allreduce(const UpdateParams& p, double* sums, double* scratch) {
    const int G = p.n_ranks;
    const int nv = 3 * p.nsteps;
    const int nw = 2 * nv;                      // words per row
    const int half = static_cast<int>(static_cast<long long>(p.tag) & 1LL);
    const unsigned int flag = static_cast<unsigned int>(static_cast<long long>(p.tag));
    // send: G * nw words, one per thread (G <= 16, nw <= 48: at most three rounds of the block)
    for (int e = threadIdx.x; e < G * nw; e += blockDim.x) {
        const int q = e / nw, k = e - q * nw;
        const double v = sums[k >> 1];
        const unsigned int bits = (k & 1) ? static_cast<unsigned int>(__double2hiint(v))
                                          : static_cast<unsigned int>(__double2loint(v));
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.peer_mbox[q]) +
                                  (static_cast<size_t>(half) * G + p.rank) * MBOX_ROW + k;
        st_word_sys(dst, static_cast<unsigned long long>(bits) | (static_cast<unsigned long long>(flag) << 32));
    }
    // receive
    unsigned int* scr = reinterpret_cast<unsigned int*>(scratch);   // [G][nw] data words
    bool ok = true;
    for (int e = threadIdx.x; e < G * nw; e += blockDim.x) {
        const int q = e / nw, k = e - q * nw;
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.peer_mbox[p.rank]) +
                                        (static_cast<size_t>(half) * G + q) * MBOX_ROW + k;
        const long long t0 = clock64();
        unsigned long long w = ld_word_sys(src);
        while (static_cast<unsigned int>(w >> 32) != flag) {
            if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; do not hang the GPU
                ok = false;
                break;
            }
            w = ld_word_sys(src);
        }
        scr[q * nw + k] = static_cast<unsigned int>(w);
    }
    if (!ok && p.error_flag != nullptr) *p.error_flag = 1;
    const int any_bad = __syncthreads_or(ok ? 0 : 1);
    if (threadIdx.x < nv) {
        double t = 0.0;
        for (int r = 0; r < G; ++r) {
            const unsigned int lo = scr[r * nw + 2 * threadIdx.x], hi = scr[r * nw + 2 * threadIdx.x + 1];
            t += __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
        }
        sums[threadIdx.x] = any_bad ? nan("") : t;
    }
    __syncthreads();
}

// The shard masses a following resample splits its offspring by: rank r's own sum w' of the last fused step, taken
// from the rows the all-reduce received.  Written AFTER the stats block has been released (nothing on the device
// waits for it) and, like the stats mirror, without a system-scope fence: 32-byte vector stores {3 sums, tag}, each a
// single aligned PCIe write; the host accepts the block when every group carries the launch's tag.
__device__ void publish_shard_norms(const UpdateParams& p, const double* scratch) {
    const int G = p.n_ranks;
    const int g = static_cast<int>(threadIdx.x) - 32;      // warp 1: thread 0 is busy publishing the stats block
    if (g < 0 || 3 * g >= G) return;
    const unsigned int* scr = reinterpret_cast<const unsigned int*>(scratch);
    const int nw = 6 * p.nsteps;
    const int k = 3 * (p.nsteps - 1);
    double v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int r = 3 * g + j;
        v[j] = 0.0;
        if (r < G) {
            const unsigned int lo = scr[r * nw + 2 * k], hi = scr[r * nw + 2 * k + 1];
            v[j] = __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
        }
    }
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.shard_norms + 4 * g), "d"(v[0]), "d"(v[1]),
                 "d"(v[2]), "d"(p.tag)
                 : "memory");
}

// MLEModel's power is applied by the one-update kernels only (fused launches keep the plain likelihood and their
// register budget; the host does not fuse updates of a model with a power, like tomography).
template <int KIND, bool BINOM, int KF, typename Row, typename Meas>
__device__ __forceinline__ double update_likelihood(const ModelView& mv, const ExpView& ev, Row row, Meas meas, int rot) {
    if constexpr (KF == 1)
        return model_likelihood<KIND, BINOM>(mv, ev, row, meas, rot);
    else
        return model_likelihood_plain<KIND, BINOM>(mv, ev, row, meas, rot);
}

// DT > 0: compile-time n_modelparams, pair processing.  DT == 0: runtime d.  KF: 1 or KF_MAX fused updates.
//
// Warp-specialised: warps 0..UPD_CONSUMER_WARPS-1 compute, the last warp's lane 0 is the TMA producer.
// full[s]  (count 1)  : producer's expect_tx + the bulk copies' complete_tx  -> consumers may read stage s
// empty[s] (count NCW): one arrive per consumer warp                        -> producer may refill stage s
// No CTA-wide barrier in the tile loop: warps drift freely across the ring.
// CHAIN: compiled for clouds whose launches may be chained (sharded clouds): flag-based dependency on the
// predecessor, ring prefetch before the stats wait, early release of the weights.  CHAIN = false is the plain
// programmatic-dependent-launch kernel (one GPU: measured 1.5 us/launch faster than running the chain-capable
// code un-chained).
template <int KIND, bool BINOM, int DT, int KF, bool CHAIN>
__global__ void __launch_bounds__(UPD_THREADS, (KF == 1) ? 2 : 3) fused_update_kernel(const __grid_constant__ UpdateParams p) {
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
    __shared__ double red[(UPD_THREADS / 32) * KF * 3];
    __shared__ double sums[3 * KF_MAX];
    __shared__ double peer_scratch[QB_MAX_RANKS * MBOX_ROW / 2];  // [ranks][words] as 32-bit data
    __shared__ unsigned int is_last;

    const int tid = threadIdx.x;
    // Programmatic dependent launch: let the NEXT launch on this stream become resident while this one runs
    // (its CTAs take free slots and wait), and do not touch anything the PREVIOUS launch wrote (stats_in, w_in)
    // before it is there.
    //   plain  : griddepcontrol.wait — the previous launch has completed and flushed.
    //   chained: the previous launch is the update with tag chain_prev_tag.  Its weights are complete once its
    //            last block has seen every block's ticket and released `data_tag`; its stats block follows a few
    //            microseconds later (final reduction, the all-reduce over the peers' mailboxes when the cloud is
    //            sharded, publish).  This launch acquires data_tag, starts streaming tiles, and only then waits
    //            for the stats tag: the predecessor's tail and the NVLink exchange hide behind the pipeline fill.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool chained = CHAIN && p.chain_prev_tag != 0.0;
    if (!chained) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    } else if (tid == 0) {
        double seen;
        do {
            asm volatile("ld.acquire.gpu.global.f64 %0, [%1];" : "=d"(seen) : "l"(p.data_tag) : "memory");
            if (seen != p.chain_prev_tag) __nanosleep(20);
        } while (seen != p.chain_prev_tag);
    }
    if constexpr (!CHAIN) {
        if (p.guard && needs_host(p.stats_in)) {
            // Speculative launch whose predecessor needs the host: do nothing, say so.  stats_out is the block of
            // the COMMITTED state the host may still fall back to, so only its SKIPPED word is touched (a guarded
            // launch queued behind this one cancels on it); the host learns through the mirror.
            if (blockIdx.x == 0 && tid == 0) {
                p.stats_out[QB_STAT_SKIPPED] = 1.0;
                if (p.mirror != nullptr) {
                    for (int j = 0; j < p.nsteps; ++j) {
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j), "d"(0.0),
                                     "d"(0.0), "d"(0.0), "d"(p.tag)
                                     : "memory");
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j + 4),
                                     "d"(0.0), "d"(0.0), "d"(p.tag), "d"(2.0)
                                     : "memory");
                    }
                }
            }
            return;
        }
    }

    const uint32_t bar0 = smem_u32(smem_raw);  // full[s] at bar0 + 8 s, empty[s] at bar0 + 64 + 8 s
    const uint32_t ring0 = smem_u32(ring);
    // the ring carries FULL tiles only; the ragged remainder (n % tile particles) is read directly by the last block
    const int ntiles = static_cast<int>(p.n / tile);
    const int my_tiles = (ntiles > static_cast<int>(blockIdx.x))
                             ? (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 0;
    const int nsteps = (KF == 1) ? 1 : p.nsteps;

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

    Acc acc[KF];
    unsigned int bad = 0u;
#pragma unroll
    for (int j = 0; j < KF; ++j) acc[j] = {0.0, 0.0};
    // Zig-zag: consecutive launches walk the slab in opposite directions.  One update streams 8(d+2) n bytes —
    // twice the L2 at n = 1e7 — so a launch that restarts at tile 0 finds nothing of the previous pass; starting
    // where the previous launch ENDED finds the tail of x and of the weights it just wrote still resident.
    const int64_t last_tile = static_cast<int64_t>(ntiles) - 1;
    auto tile_first = [&](int i) -> int64_t {
        const int64_t ti = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
        return (p.reverse ? last_tile - ti : ti) * tile;
    };

    if constexpr (CHAIN) {
        // ===== producer state (lane 0 of the last warp streams my tiles into the ring) =====
        int prod_s = 0;
        uint32_t prod_phase = 0;
        // L2 residency: w_in is dead once read (the next launch overwrites it), x and the new weights are what
        // the next launch reads first (it walks the slab in the opposite direction)
        uint64_t pol_w = (p.l2hint & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
        uint64_t pol_x = (p.l2hint & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
        const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
        const int64_t pin_end = static_cast<int64_t>(p.pin_tiles) * tile;
        auto produce = [&](int i_from, int i_to) {
            for (int i = i_from; i < i_to; ++i) {
                const int64_t first = tile_first(i);
                if (p.l2hint & 8) pol_w = pol_x = (first < pin_end) ? pol_keep : pol_stream;
                mbar_wait(bar0 + 64 + 8 * prod_s, prod_phase ^ 1u);  // passes at once the first time round the ring
                const uint32_t dst = ring0 + static_cast<uint32_t>(prod_s) * stage_bytes;
                mbar_expect_tx(bar0 + 8 * prod_s, stage_bytes);
                tma_load_1d_hint(dst, p.x + first * d, x_bytes, bar0 + 8 * prod_s, pol_x);
                tma_load_1d_hint(dst + x_bytes, p.w_in + first, w_bytes, bar0 + 8 * prod_s, pol_w);
                if (++prod_s == UPD_STAGES) {
                    prod_s = 0;
                    prod_phase ^= 1u;
                }
            }
        };
        // Fill the ring before anything else: these copies need the predecessor's weights (acquired above), not its stats.
        const int npre = (my_tiles < UPD_STAGES) ? my_tiles : UPD_STAGES;
        if (tid == NCT) {
            if (chained) asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy acquire -> async-proxy reads
            produce(0, npre);
        }
        if (chained && tid == 0) {  // now the predecessor's stats block (its last block is still reducing / exchanging)
            double seen;
            do {
                asm volatile("ld.acquire.gpu.global.f64 %0, [%1];" : "=d"(seen) : "l"(p.stats_in + QB_STAT_TAG) : "memory");
                if (seen != p.chain_prev_tag) __nanosleep(20);
            } while (seen != p.chain_prev_tag);
        }
        __syncthreads();
        if (p.guard && needs_host(p.stats_in)) {
            // Speculative launch whose predecessor needs the host: do nothing, say so.  stats_out is the block of
            // the COMMITTED state the host may still fall back to, so only its SKIPPED word is touched (a guarded
            // launch queued behind this one cancels on it); the host learns through the mirror.
            if (tid == 0) {
                for (int s = 0; s < npre; ++s) mbar_wait(bar0 + 8 * s, 0u);  // let the prefetched copies land before exit
                if (blockIdx.x == 0) {
                    p.stats_out[QB_STAT_SKIPPED] = 1.0;
                    __threadfence();
                    *reinterpret_cast<volatile double*>(p.stats_out + QB_STAT_TAG) = p.tag;
                    if (p.mirror != nullptr) {
                        for (int j = 0; j < p.nsteps; ++j) {
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j), "d"(0.0),
                                         "d"(0.0), "d"(0.0), "d"(p.tag)
                                         : "memory");
                            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p.mirror + 8 * j + 4),
                                         "d"(0.0), "d"(0.0), "d"(p.tag), "d"(2.0)
                                         : "memory");
                        }
                    }
                    // a chained successor of THIS launch must not wait for weights that will never come
                    if (p.data_tag != nullptr) {
                        __threadfence();
                        *reinterpret_cast<volatile double*>(p.data_tag) = p.tag;
                    }
                }
            }
            __syncthreads();
            return;
        }

        if (tid == NCT) produce(npre, my_tiles);   // the producer lane does the rest of its job here
    }

    if (tid >= NCT) {
        if constexpr (!CHAIN) {
            // ===== producer warp: one lane streams my tiles into the ring =====
            if (tid == NCT) {
                int s = 0;
                uint32_t phase = 0;
                // L2 residency: w_in is dead once read (the next launch overwrites it), x and the new weights are
                // what the next launch reads first (it walks the slab in the opposite direction)
                uint64_t pol_w = (p.l2hint & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
                uint64_t pol_x = (p.l2hint & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
                const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
                const int64_t pin_end = static_cast<int64_t>(p.pin_tiles) * tile;
                for (int i = 0; i < my_tiles; ++i) {
                    const int64_t first = tile_first(i);
                    if (p.l2hint & 8) pol_w = pol_x = (first < pin_end) ? pol_keep : pol_stream;
                    mbar_wait(bar0 + 64 + 8 * s, phase ^ 1u);  // passes at once the first time round the ring
                    const uint32_t dst = ring0 + static_cast<uint32_t>(s) * stage_bytes;
                    mbar_expect_tx(bar0 + 8 * s, stage_bytes);
                    tma_load_1d_hint(dst, p.x + first * d, x_bytes, bar0 + 8 * s, pol_x);
                    tma_load_1d_hint(dst + x_bytes, p.w_in + first, w_bytes, bar0 + 8 * s, pol_w);
                    if (++s == UPD_STAGES) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===== consumer warps =====
        const double inv_norm = CHAIN ? __ldcg(p.stats_in + QB_STAT_INV_NORM) : p.stats_in[QB_STAT_INV_NORM];
        const ModelView mv = p.mv;
        const int lane = tid & 31;
        auto meas = [&](int c) { return meas_s[c]; };
        int s = 0;
        uint32_t phase = 0;
        const bool keep_out = (p.l2hint & 4) != 0;
        uint64_t pol_o = keep_out ? l2_policy_evict_last() : l2_policy_evict_normal();
        const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
        const int64_t pin_end = static_cast<int64_t>(p.pin_tiles) * tile;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t first = tile_first(i);
            if (p.l2hint & 8) pol_o = (first < pin_end) ? pol_keep : pol_stream;
            double* wo = p.w_out + first;
            {
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
                        double w0 = wp.x * inv_norm;  // the lazily applied normalisation of the previous launch
                        double w1 = wp.y * inv_norm;
#pragma unroll
                        for (int k = 0; k < KF; ++k) {
                            if (KF == 1 || k < nsteps) {
                                w0 = w0 * update_likelihood<KIND, BINOM, KF>(mv, p.ev[k], row0, meas, 0);  // smc.py:354
                                w1 = w1 * update_likelihood<KIND, BINOM, KF>(mv, p.ev[k], row1, meas, 0);
                                accumulate(acc[k], bad, k, w0);
                                accumulate(acc[k], bad, k, w1);
                            }
                        }
                        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(
                                         wo + 2 * j),
                                     "d"(w0), "d"(w1), "l"(pol_o)
                                     : "memory");
                    }
                } else {
                    for (int j = tid; j < tile; j += NCT) {
                        const double* xr = xs + static_cast<size_t>(j) * d;
                        auto row = [&](int c) { return xr[c]; };
                        double wv = ws[j] * inv_norm;
#pragma unroll
                        for (int k = 0; k < KF; ++k) {
                            if (KF == 1 || k < nsteps) {
                                wv = wv * update_likelihood<KIND, BINOM, KF>(mv, p.ev[k], row, meas, lane);
                                accumulate(acc[k], bad, k, wv);
                            }
                        }
                        stg_stream(wo + j, wv);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 64 + 8 * s);  // this warp is done with stage s
            }
            if (++s == UPD_STAGES) {
                s = 0;
                phase ^= 1u;
            }
        }
        if (blockIdx.x == gridDim.x - 1) {  // ragged remainder: straight from global memory
            const int64_t first = static_cast<int64_t>(ntiles) * tile;
            const int cnt = static_cast<int>(p.n - first);
            for (int j = tid; j < cnt; j += NCT) {
                const double* xr = p.x + (first + j) * d;
                auto row = [&](int c) { return xr[c]; };
                double wv = (CHAIN ? __ldcg(p.w_in + first + j) : p.w_in[first + j]) * inv_norm;
#pragma unroll
                for (int k = 0; k < KF; ++k) {
                    if (KF == 1 || k < nsteps) {
                        wv = wv * update_likelihood<KIND, BINOM, KF>(mv, p.ev[k], row, meas, 0);
                        accumulate(acc[k], bad, k, wv);
                    }
                }
                if (CHAIN)
                    __stcg(p.w_out + first + j, wv);
                else
                    p.w_out[first + j] = wv;
            }
        }
    }

    block_reduce_steps<KF>(acc, bad, red, sums);
    if (tid < 3 * KF) p.partials[static_cast<size_t>(blockIdx.x) * 3 * KF + tid] = sums[tid];
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(p.ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        // every block's weights are written (each fenced before its ticket): release them to a chained successor,
        // which then streams tiles while this block finishes the reduction (and the NVLink exchange)
        __threadfence();
        if (CHAIN && tid == 0) {
            *p.ticket = 0u;  // no block of this launch touches it again; the successor's blocks come much later
            *reinterpret_cast<volatile double*>(p.data_tag) = p.tag;  // ordered behind the fence above
        }
        // deterministic final reduction of the per-block partials (fixed lane/block order)
        constexpr int NV = 3 * KF;
        const int lane = tid & 31, wid = tid >> 5, nw = UPD_THREADS / 32;
        if (tid < 3 * KF_MAX) sums[tid] = 0.0;
        __syncthreads();
        for (int v = wid; v < NV; v += nw) {  // one warp per value
            double t = 0.0;
            for (int b = lane; b < static_cast<int>(gridDim.x); b += 32) {
                const double* pp = p.partials + static_cast<size_t>(b) * NV + v;
                t += CHAIN ? __ldcg(pp) : *pp;
            }
            t = warp_sum(t);
            if (lane == 0) sums[v] = t;
        }
        __syncthreads();
        if (p.n_ranks > 1) peer_allreduce(p, sums, peer_scratch);
        if (tid == 0) {
            publish(p, sums);
            if (!CHAIN) *p.ticket = 0u;  // ready for the next launch on this stream
        } else if (p.n_ranks > 1 && p.shard_norms != nullptr && tid >= 32) {
            publish_shard_norms(p, peer_scratch);
        }
    }
}

// Tile size: 16-24 KB per stage; a multiple of 2 * consumer threads for the pair kernels,
// of 16 particles otherwise, so every bulk copy is 128-B granular.
static int choose_tile(int d) {
    if (d == 1) return 1024;           // 16 KB / stage
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

template <int KF, bool CH>
static update_kernel_t pick_update_kernel_k(const qb_model& m) {
    switch (m.kind) {
        case QB_MODEL_PRECESSION:
            return m.binomial ? fused_update_kernel<QB_MODEL_PRECESSION, true, 1, KF, CH>
                              : fused_update_kernel<QB_MODEL_PRECESSION, false, 1, KF, CH>;
        case QB_MODEL_RB:
            if (m.interleaved)
                return m.binomial ? fused_update_kernel<QB_MODEL_RB, true, 4, KF, CH>
                                  : fused_update_kernel<QB_MODEL_RB, false, 4, KF, CH>;
            return m.binomial ? fused_update_kernel<QB_MODEL_RB, true, 3, KF, CH>
                              : fused_update_kernel<QB_MODEL_RB, false, 3, KF, CH>;
        case QB_MODEL_COIN:
            return m.binomial ? fused_update_kernel<QB_MODEL_COIN, true, 1, KF, CH>
                              : fused_update_kernel<QB_MODEL_COIN, false, 1, KF, CH>;
    }
    return nullptr;
}

// a decorator appended parameters the likelihood ignores: the row pitch is a run-time value (generic-d kernel, K = 1)
static update_kernel_t pick_update_kernel_extra(const qb_model& m) {
    switch (m.kind) {
        case QB_MODEL_PRECESSION:
            return m.binomial ? fused_update_kernel<QB_MODEL_PRECESSION, true, 0, 1, false>
                              : fused_update_kernel<QB_MODEL_PRECESSION, false, 0, 1, false>;
        case QB_MODEL_RB:
            return m.binomial ? fused_update_kernel<QB_MODEL_RB, true, 0, 1, false>
                              : fused_update_kernel<QB_MODEL_RB, false, 0, 1, false>;
        case QB_MODEL_COIN:
            return m.binomial ? fused_update_kernel<QB_MODEL_COIN, true, 0, 1, false>
                              : fused_update_kernel<QB_MODEL_COIN, false, 0, 1, false>;
    }
    return nullptr;
}

template <bool CH>
static update_kernel_t pick_update_kernel_c(const qb_model& m, int nsteps) {
    if (m.d_extra > 0 && m.kind != QB_MODEL_TOMOGRAPHY) return (CH || nsteps != 1) ? nullptr : pick_update_kernel_extra(m);
    if (m.kind == QB_MODEL_TOMOGRAPHY)  // per-step measurement vectors do not fit the launch parameters: K = 1
        return m.binomial ? fused_update_kernel<QB_MODEL_TOMOGRAPHY, true, 0, 1, CH>
                          : fused_update_kernel<QB_MODEL_TOMOGRAPHY, false, 0, 1, CH>;
    return (nsteps == 1) ? pick_update_kernel_k<1, CH>(m) : pick_update_kernel_k<KF_MAX, CH>(m);
}

static update_kernel_t pick_update_kernel(const qb_model& m, int nsteps, bool chain) {
    return chain ? pick_update_kernel_c<true>(m, nsteps) : pick_update_kernel_c<false>(m, nsteps);
}

int validate_model(const qb_model* m) {
    QB_REQUIRE(m != nullptr, QB_ERR_INVALID_ARGUMENT, "model is NULL");
    QB_REQUIRE(m->d >= 1 && m->d <= QB_MAX_D, QB_ERR_INVALID_ARGUMENT, "n_modelparams %d outside [1, %d]", m->d,
               QB_MAX_D);
    QB_REQUIRE(m->d_extra >= 0 && m->d_extra < m->d && m->extra_rule >= 0 && m->extra_rule <= 2,
               QB_ERR_INVALID_ARGUMENT, "bad d_extra %d / extra_rule %d", m->d_extra, m->extra_rule);
    const int db = m->d - m->d_extra;   // parameters of the likelihood itself
    switch (m->kind) {
        case QB_MODEL_PRECESSION:
            QB_REQUIRE(db == 1, QB_ERR_UNSUPPORTED_MODEL, "precession model has 1 model parameter, got %d", db);
            break;
        case QB_MODEL_RB:
            QB_REQUIRE(db == (m->interleaved ? 4 : 3), QB_ERR_UNSUPPORTED_MODEL,
                       "RB model needs %d model parameters, got %d", m->interleaved ? 4 : 3, db);
            break;
        case QB_MODEL_TOMOGRAPHY:
            break;
        case QB_MODEL_COIN:
            QB_REQUIRE(db == 1, QB_ERR_UNSUPPORTED_MODEL, "coin model has 1 model parameter, got %d", db);
            break;
        default:
            set_error("unknown model kind %d (no CPU fallback exists)", m->kind);
            return QB_ERR_UNSUPPORTED_MODEL;
    }
    return QB_OK;
}

static int update_grid_limit(update_kernel_t k, size_t smem, int nsteps) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, UPD_THREADS, smem) != cudaSuccess) return -1;
    if (per_sm < 1) per_sm = 1;
    // Two CTAs per SM per launch: leaves room for the next (programmatically dependent) launch to become
    // resident while this one runs; measured faster than filling the SM with one launch (r1: 43.6 vs 46.6 us).
    // The fused-batch variant is fp64-latency bound instead: it takes every CTA slot it can get.
    int cap = (nsteps == 1) ? 2 : 8;
    const char* e = getenv("QB_UPD_CTAS_PER_SM");  // experiment knob
    if (e && atoi(e) > 0) cap = atoi(e);
    if (per_sm > cap) per_sm = cap;
    return per_sm * sm_count();
}

struct GridCacheEntry {
    update_kernel_t k;
    int dev;
    int limit;
};
static GridCacheEntry g_grid_cache[64];
static int g_grid_cache_n = 0;

static int cached_grid_limit(update_kernel_t k, size_t smem, int nsteps) {
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < g_grid_cache_n; ++i)
        if (g_grid_cache[i].k == k && g_grid_cache[i].dev == dev) return g_grid_cache[i].limit;
    const int limit = update_grid_limit(k, smem, nsteps);
    if (limit > 0 && g_grid_cache_n < 64) g_grid_cache[g_grid_cache_n++] = {k, dev, limit};
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
    // per-block partials (3 * QB_MAX_FUSE doubles) for up to 32 blocks/SM on up to 256 SMs + the ticket
    return static_cast<size_t>(32) * 256 * 3 * QB_MAX_FUSE * sizeof(double) + 256;
}

extern "C" int qb_fused_update_multi(const qb_model* model, const qb_expparams* eps, const int64_t* outcomes,
                                     int32_t nsteps, uint32_t resample_mask, const double* d_x, int64_t n,
                                     const double* d_w_in, double* d_w_out, const double* d_stats_in,
                                     double* d_stats_out, double* d_step_stats, const qb_update_ctl* ctl, void* d_ws,
                                     size_t ws_bytes, void* stream) {
    int rc = validate_model(model);
    if (rc != QB_OK) return rc;
    QB_REQUIRE(eps && outcomes && d_x && d_w_in && d_w_out && d_stats_in && d_stats_out && d_ws,
               QB_ERR_INVALID_ARGUMENT, "qb_fused_update: NULL pointer argument");
    QB_REQUIRE(n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: n must be >= 1, got %lld", (long long)n);
    QB_REQUIRE(nsteps >= 1 && nsteps <= QB_MAX_FUSE, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: 1..%d updates per launch, got %d", QB_MAX_FUSE, nsteps);
    QB_REQUIRE(nsteps == 1 || model->kind != QB_MODEL_TOMOGRAPHY, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: tomography updates are not fused (one measurement vector per launch)");
    QB_REQUIRE(nsteps == 1 || model->likelihood_power == 0.0 || model->likelihood_power == 1.0, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: updates of a model with a likelihood power (MLEModel) are not fused");
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
    p.step_stats = d_step_stats;
    p.ticket = reinterpret_cast<unsigned int*>(d_ws);
    p.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
    p.n = n;
    p.d = model->d;
    p.tile = choose_tile(model->d);
    p.nsteps = nsteps;
    p.resample_mask = resample_mask;
    p.mirror = ctl ? ctl->h_mirror : nullptr;
    p.tag = ctl ? ctl->tag : 0.0;
    p.zero_weight_thresh = ctl ? ctl->zero_weight_thresh : 0.0;
    p.resample_below = ctl ? ctl->resample_below : 0.0;
    p.guard = ctl ? ctl->guard : 0;
    {
        // experiment knobs (defaults are the measured best): QB_UPD_ZIGZAG=0 disables the alternating direction,
        // QB_UPD_L2HINT=<bits> selects the L2 eviction hints
        static int zig = -1, hint = -1, pin_mb = 72;
        if (zig < 0) {
            const char* e1 = getenv("QB_UPD_ZIGZAG");
            const char* e2 = getenv("QB_UPD_L2HINT");
            const char* e3 = getenv("QB_UPD_PIN_MB");
            zig = e1 ? atoi(e1) : 1;
            hint = e2 ? atoi(e2) : 5;
            if (e3) pin_mb = atoi(e3);
        }
        // the ping-pong weight buffers swap roles every launch: the direction follows which one is the source
        p.reverse = (zig && reinterpret_cast<uintptr_t>(d_w_in) > reinterpret_cast<uintptr_t>(d_w_out)) ? 1 : 0;
        p.l2hint = hint;
        // partition mode: pin_mb MB of L2 shared by the head of x and of BOTH weight buffers
        const double per_tile = static_cast<double>(choose_tile(model->d)) * 8.0 * (model->d + 2);
        p.pin_tiles = static_cast<int32_t>(static_cast<double>(pin_mb) * 1e6 / per_tile);
    }
    {
        static int chain_on = -1;
        if (chain_on < 0) {
            const char* e = getenv("QB_UPD_CHAIN");  // 0 = never, 1 = sharded clouds (default), 2 = always
            chain_on = e ? atoi(e) : 1;
        }
        // default: chain the launches of a SHARDED cloud (the all-reduce then hides behind the successor's
        // pipeline fill: 44.1 -> 40.9 us/launch at 2 GPUs); on one GPU programmatic dependent launch alone is as fast
        const bool want = ctl && (chain_on == 1 ? (ctl->n_ranks > 1) : (chain_on == 2));
        p.chain_prev_tag = want ? ctl->chain_prev_tag : 0.0;
        p.chain_capable = want ? 1 : 0;
        p.pad2 = 0;
        p.data_tag = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_ws) + 64);
    }
    p.n_ranks = ctl ? ctl->n_ranks : 0;
    p.rank = ctl ? ctl->rank : 0;
    p.error_flag = ctl ? ctl->d_error_flag : nullptr;
    p.shard_norms = (ctl && ctl->n_ranks > 1) ? ctl->h_shard_norms : nullptr;
    QB_REQUIRE(p.mirror == nullptr || (reinterpret_cast<uintptr_t>(p.mirror) & 31) == 0, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: the host mirror must be 32-byte aligned");
    QB_REQUIRE(p.n_ranks <= QB_MAX_RANKS, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: at most %d ranks", QB_MAX_RANKS);
    QB_REQUIRE(p.shard_norms == nullptr || (reinterpret_cast<uintptr_t>(p.shard_norms) & 31) == 0, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: h_shard_norms must be 32-byte aligned");
    for (int r = 0; r < QB_MAX_RANKS; ++r) p.peer_mbox[r] = (ctl && r < p.n_ranks) ? ctl->d_peer_mailbox[r] : nullptr;
    if (p.n_ranks > 1) {
        QB_REQUIRE(p.rank >= 0 && p.rank < p.n_ranks, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: bad rank");
        for (int r = 0; r < p.n_ranks; ++r)
            QB_REQUIRE(p.peer_mbox[r] != nullptr, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: NULL peer mailbox %d", r);
    }
    QB_REQUIRE(d_stats_in != d_stats_out || !p.guard, QB_ERR_INVALID_ARGUMENT,
               "qb_fused_update: a guarded update needs distinct stats_in / stats_out blocks");
    p.mv = make_model_view(*model);
    for (int j = 0; j < KF_MAX; ++j) {
        const int jj = (j < nsteps) ? j : 0;
        p.ev[j] = make_exp_view(*model, eps[jj], outcomes[jj]);
    }
    for (int c = 0; c < QB_MAX_D; ++c) p.meas[c] = (c < model->d) ? eps[0].meas[c] : 0.0;

    update_kernel_t k = pick_update_kernel(*model, nsteps, p.chain_capable != 0);
    QB_REQUIRE(k != nullptr, QB_ERR_UNSUPPORTED_MODEL,
               "qb_fused_update: a model with decorator parameters (d_extra > 0) takes one update per launch on an "
               "unsharded cloud");
    const size_t smem = update_smem_bytes(model->d);
    const int limit = cached_grid_limit(k, smem, nsteps);
    QB_REQUIRE(limit > 0, QB_ERR_CUDA, "qb_fused_update: occupancy query failed: %s",
               cudaGetErrorString(cudaGetLastError()));
    const int64_t ntiles = (n + p.tile - 1) / p.tile;
    int grid = static_cast<int>(ntiles < limit ? ntiles : limit);
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

extern "C" int qb_fused_update(const qb_model* model, const qb_expparams* ep, int64_t outcome, const double* d_x,
                               int64_t n, const double* d_w_in, double* d_w_out, const double* d_stats_in,
                               double* d_stats_out, const qb_update_ctl* ctl, void* d_ws, size_t ws_bytes,
                               void* stream) {
    QB_REQUIRE(ep != nullptr, QB_ERR_INVALID_ARGUMENT, "qb_fused_update: NULL pointer argument");
    const uint32_t mask = (ctl && ctl->check_resample) ? 1u : 0u;
    return qb_fused_update_multi(model, ep, &outcome, 1, mask, d_x, n, d_w_in, d_w_out, d_stats_in, d_stats_out,
                                 nullptr, ctl, d_ws, ws_bytes, stream);
}

extern "C" int qb_weights_min(const double* d_w, int64_t n, double* d_out, void* stream) {
    QB_REQUIRE(d_w && d_out && n >= 1, QB_ERR_INVALID_ARGUMENT, "qb_weights_min: bad arguments");
    weights_min_kernel<<<1, 256, 0, as_stream(stream)>>>(d_w, n, d_out);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
