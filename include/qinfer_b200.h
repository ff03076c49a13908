/* qinfer_b200.h — C ABI of the B200-native SMC particle-filter hot path.
 *
 * This is the drop-in boundary for QInfer's Bayes-update + Liu-West resample
 * path.  The reference (QInfer/python-qinfer @ 8170c84) is pure
 * Python/NumPy and has no FFI of its own; these entry points are what a ctypes
 * binding of that path binds (INTEGRATION.md shows the stub), one per reference
 * function listed in SURVEY §8(a).  File:line citations are relative to
 * /root/reference/src/qinfer/.
 *
 * Conventions
 *   - Plain pointers and sizes only; no framework types.  Every `*_dev` / `d_`
 *     pointer is CUDA device memory owned by the caller (the Python host
 *     allocates it with torch and passes `data_ptr()`); `stream` is a
 *     `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *   - All particle data is float64.  Particle locations are row-major
 *     `x[n][d]` exactly like `SMCUpdater.particle_locations` (smc.py:292-298).
 *   - Weights are kept UNNORMALISED on the device together with a 16-double
 *     `stats` block (layout below); the normalised weight of particle i is
 *     `w[i] * stats[QB_STAT_INV_NORM]`.  This defers the division of
 *     smc.py:373 into the next pass over the weights (8(d+2) B/particle-update).
 *   - Every function returns QB_OK (0) or a negative QB_ERR_* code;
 *     `qb_last_error()` returns a thread-local message.  Launches are
 *     asynchronous on `stream`; nothing here synchronises unless stated.
 *   - No entry point allocates device memory; workspaces are caller-provided
 *     and sized by the `*_workspace_bytes` queries.  One workspace may be shared
 *     by all calls issued on one stream; it must be ZERO-INITIALISED once before
 *     first use (its first 256 bytes hold the fused-update kernel's self-resetting
 *     last-block ticket, which no other entry point touches).
 */
#ifndef QINFER_B200_H
#define QINFER_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_ABI_VERSION 4 /* 4: qb_update_ctl.h_shard_norms, qb_cdf_chained, qb_lw_binned_shard_consts; 2: qb_model.likelihood_power, qb_update_ctl.chain_prev_tag, flagged-word mailboxes; 3: qb_lw_binned_*,
                            qb_model.d_extra / extra_rule, qb_walk_step, qb_poison_likelihood, qb_tomo_canonicalize*_ld,
                            qb_weights_entropy, qb_weight_mass_hist, qb_weights_select */

#define QB_OK 0
#define QB_ERR_INVALID_ARGUMENT (-1)
#define QB_ERR_UNSUPPORTED_MODEL (-2)
#define QB_ERR_CUDA (-3)
#define QB_ERR_WORKSPACE (-4)

#define QB_MAX_D 64 /* max n_modelparams handled by the staged kernels (3-qubit tomography) */
#define QB_MAX_RANKS 16 /* GPUs of one NVLink domain that may share a particle cloud */
#define QB_MAX_FUSE 8   /* consecutive updates one launch can fuse (qb_fused_update_multi) */
#define QB_MAILBOX_ROW 64 /* 8-byte slots per peer-mailbox row: 2 per sum (3 * QB_MAX_FUSE sums), each {32 data bits, 32-bit flag} */

/* ---- model plugin descriptor ------------------------------------------- */
/* Which built-in likelihood the kernels evaluate. */
enum qb_model_kind {
    QB_MODEL_PRECESSION = 1, /* SimpleInversionModel / SimplePrecessionModel  test_models.py:123-143,188-197 */
    QB_MODEL_RB = 2,         /* RandomizedBenchmarkingModel                   rb.py:149-195                  */
    QB_MODEL_TOMOGRAPHY = 3, /* tomography.TomographyModel                    tomography/models.py:143-226   */
    QB_MODEL_COIN = 4        /* CoinModel (pr0 = p; the model of the reference's risk / information-gain
                                known-answer tests, tests/test_metrics.py)    test_models.py:262-326         */
};

typedef struct qb_model {
    int32_t kind;         /* enum qb_model_kind */
    int32_t d;            /* n_modelparams (abstract_model.py:96-104) */
    int32_t binomial;     /* 1: wrapped in BinomialModel (derived_models.py:222-360) */
    int32_t interleaved;  /* RB only: `_il` (rb.py:114-115) */
    double min_freq;      /* precession only: `_min_freq` (test_models.py:78-80,109-110) */
    double likelihood_power; /* MLEModel (derived_models.py:681-703): every likelihood is raised to this power
                                after the (binomial) model evaluated it; 0 or 1 = plain model */
    int32_t d_extra;      /* trailing model parameters the likelihood ignores (d counts them): the learned step scales
                             of GaussianRandomWalkModel (derived_models.py:811-818, 894-896), the diffusion rate of
                             DiffusiveTomographyModel (tomography/models.py:229-256) */
    int32_t extra_rule;   /* their validity (are_models_valid): 0 none, 1 all >= 0 (derived_models.py:883-892),
                             2 the last one > 0 (tomography/models.py:245-249) */
    int32_t fast_math;    /* 0 (default): the reference's operation sequence — pow() for p ** m (rb.py:193), SciPy's
                             exp(logC + k log p + (n-k) log1p(-p)) for the binomial pmf (utils.py:106-111).
                             1: integer powers by squaring — p ** m (m is an unsigned integer in the reference) and, for
                             n_meas <= 56, C(n,k) p^k (1-p)^(n-k) with the exact binomial coefficient: no pow / log /
                             exp per particle.  Relative deviation from the strict path <= 1e-13 (each of the <= 2 log2
                             multiplications rounds once); north_star's 1e-6 on mean / covariance is unaffected. */
    int32_t reserved0;
} qb_model;

/* One experiment record (one element of the `expparams` array handed to
 * Model.likelihood, abstract_model.py:443-468). */
typedef struct qb_expparams {
    double t;            /* precession: evolution time 't'                       */
    double w_;           /* inversion model: reference frequency 'w_' (0 for SimplePrecessionModel) */
    int64_t m;           /* RB: sequence length 'm' (uint in the reference)      */
    int32_t reference;   /* interleaved RB: 'reference' flag                     */
    int32_t reserved;
    int64_t n_meas;      /* BinomialModel: 'n_meas'                              */
    double meas[QB_MAX_D]; /* tomography: 'meas' coefficients (first d used)     */
} qb_expparams;

/* ---- device-side stats block (16 doubles) ------------------------------- */
#define QB_STAT_NORM 0      /* sum_i w'_i of the last update  = normalization_record entry (smc.py:357,444) */
#define QB_STAT_SUMSQ 1     /* sum_i w'_i^2                   -> n_ess = norm^2 / sumsq (distributions.py:299-307) */
#define QB_STAT_MIN 2       /* min_i w_i: filled by qb_weights_restat/clip; NaN after qb_fused_update (see qb_weights_min) */
#define QB_STAT_NBAD 3      /* number of NaN or negative w'_i */
#define QB_STAT_INV_NORM 4  /* 1/norm, or 1 if |norm| < eps (smc.py:369-373); applied lazily by the next kernel */
#define QB_STAT_NESS 5      /* n_ess = norm^2 / sumsq (1 / sumsq when the |norm| < eps guard applies) */
#define QB_STAT_TAG 6       /* caller-chosen tag of the launch that wrote this block (qb_update_ctl.tag) */
#define QB_STAT_SKIPPED 7   /* 1 if a guarded qb_fused_update cancelled itself (see qb_update_ctl) */
#define QB_STAT_ATTN 8      /* 1-based index of the first fused step that needs the host (clip, zero-weight policy,
                               resample trigger), 0 if none; a guarded successor cancels itself when it is set */
#define QB_STAT_COUNT 16

/* ---- library / device --------------------------------------------------- */
int qb_abi_version(void);
const char* qb_last_error(void);
/* Number of SMs of the current device (grid sizing); negative on error. */
int qb_device_sm_count(void);
/* sizeof(qb_model), sizeof(qb_expparams), sizeof(qb_update_ctl) as compiled — lets a binding check its layouts. */
void qb_struct_sizes(int32_t out[3]);

/* ---- weights utilities -------------------------------------------------- */
/* w[i] = 1/n, stats = {norm 1, sumsq 1/n, min 1/n, nbad 0, inv 1, ness n}.
 * smc.py:307 (reset) and resamplers.py:390-392 (post-resample weights). */
int qb_weights_set_uniform(double* d_w, int64_t n, double* d_stats, void* stream);
/* Same for one slab of a sharded cloud: n_local weights of 1/n_global, stats describe the GLOBAL cloud. */
int qb_weights_set_uniform_global(double* d_w, int64_t n_local, int64_t n_global, double* d_stats, void* stream);
/* out[i] = w[i] * stats[INV_NORM]  — materialises `particle_weights` for the host. */
int qb_weights_normalized(const double* d_w, int64_t n, const double* d_stats, double* d_out, void* stream);
/* Re-derive stats from weights the host assigned (`particle_weights = ...`):
 * stats from sum / sum of squares of w as given (no normalisation applied). */
int qb_weights_restat(const double* d_w, int64_t n, double* d_stats, double* d_ws, size_t ws_bytes, void* stream);
/* In-place clip of the NORMALISED weights to [0,1] (smc.py:418) and restat;
 * afterwards w holds normalised, clipped weights and stats[INV_NORM] = 1. */
int qb_weights_clip(double* d_w, int64_t n, double* d_stats, double* d_ws, size_t ws_bytes, void* stream);

/* *d_out = min_i w[i] (unnormalised).  Only the warning text of smc.py:417 needs it, so the
 * fused kernel does not track it: stats[QB_STAT_MIN] is NaN after qb_fused_update. */
int qb_weights_min(const double* d_w, int64_t n, double* d_out, void* stream);

/* ---- fused Bayes update (the hot kernel) --------------------------------- */
size_t qb_update_workspace_bytes(int64_t n, int32_t d);

/* Optional launch control.  The reference decides on the host, after every update, whether the
 * weights need clipping (smc.py:416), the zero-weight policy applies (smc.py:423-436) or a resample is
 * due (smc.py:275).  Those are rare, so the host may launch the next update BEFORE it has seen the result
 * of this one: every launch evaluates the three tests on its own sums and records the outcome in
 * stats_out[QB_STAT_ATTN]; a launch with `guard` set first inspects stats_in and, if its predecessor needs
 * the host (or cancelled itself), cancels itself too (SKIPPED = 1, no weight touched).
 * `h_mirror` is a device-accessible, 32-byte aligned pinned host block of 8 doubles per fused step receiving
 *   { S_j = sum w_j, Q_j = sum w_j^2, #bad_j, TAG | normalisation record_j, n_ess_j, TAG, attention_j + 2*skipped }
 * as two 32-byte stores, each carrying the tag: the host polls plain memory and accepts a block when both
 * tags equal the launch's tag — no copy + synchronise per update and no system fence in the kernel. */
typedef struct qb_update_ctl {
    double* h_mirror;          /* may be NULL */
    double tag;
    double zero_weight_thresh; /* smc.py:171-175 */
    double resample_below;     /* n_particles * resample_thresh (smc.py:275) */
    int32_t guard;             /* 1: cancel if stats_in says the predecessor needs the host */
    int32_t check_resample;    /* qb_fused_update only: this update is followed by an n_ess check */
    /* Chained launch: if the call issued IMMEDIATELY before this one on the same stream and workspace was the
     * qb_fused_update[_multi] with this tag, and this call's w_in / stats_in are that call's w_out / stats_out, pass
     * the tag here (else 0).  The kernel then depends on its predecessor through two flags in device memory (weights
     * complete; stats published) instead of waiting for it to drain, so the predecessor's final reduction — and,
     * for a sharded cloud, its all-reduce over NVLink — overlaps with this launch's pipeline fill. */
    double chain_prev_tag;
    /* Sharded cloud: with n_ranks > 1 the kernel all-reduces (sum w', sum w'^2, #bad) of every
     * fused step over the peers' mailboxes (qb_mailbox_create / qb_ipc_*) before publishing, so stats_out
     * holds the GLOBAL normalisation and n_ess on every rank, bit-identical, with no NCCL call and no extra
     * launch.  All ranks must issue the same sequence of launches with the same tags. */
    int32_t n_ranks, rank;
    double* d_peer_mailbox[QB_MAX_RANKS]; /* [r] = rank r's mailbox as mapped in THIS process */
    int32_t* d_error_flag;     /* optional device int, set to 1 if a peer never answered */
    /* Optional, n_ranks > 1: device-accessible pinned host block of 4 * ceil(QB_MAX_RANKS / 3) doubles, 32-byte
     * aligned.  After releasing its stats block the launch stores every rank's own sum w' of its LAST fused step —
     * the shard masses a following resample splits its offspring by, known without another pass or collective — as
     * groups {sum[3g], sum[3g+1], sum[3g+2], tag}, one 32-byte store each (no system fence): valid when every group
     * that holds a rank carries the launch tag. */
    double* h_shard_norms;
} qb_update_ctl;
/* One launch: w_out[i] = (w_in[i] * stats_in[INV_NORM]) * L(outcome | x_i; ep)
 * fused with the block+warp reductions for sum, sum of squares, min and the
 * bad-weight count, finished deterministically by the last block into
 * stats_out (may alias stats_in).  Replaces SMCUpdater.hypothetical_update +
 * the weight bookkeeping of SMCUpdater.update (smc.py:324-386, 413-453) and
 * Model.likelihood for the built-in models.  w_out may alias w_in.
 * `outcome`: the datum (0/1 label, or the count k under BinomialModel). */
int qb_fused_update(const qb_model* model, const qb_expparams* ep, int64_t outcome,
                    const double* d_x, int64_t n,
                    const double* d_w_in, double* d_w_out,
                    const double* d_stats_in, double* d_stats_out,
                    const qb_update_ctl* ctl /* may be NULL */,
                    void* d_ws, size_t ws_bytes, void* stream);

/* The same launch for K = nsteps <= QB_MAX_FUSE CONSECUTIVE updates (batch_update, smc.py:459-487): the K
 * likelihoods are applied while the particle sits in registers, so the launch still moves 8(d+2) bytes per
 * particle.  eps / outcomes: HOST arrays of K records; bit j of resample_mask: step j is followed by an n_ess
 * check.  Per-step sums are published (mirror, and d_step_stats = K x 8 doubles if not NULL) so the host can
 * replay each step's record / n_ess / policy; if step j < K needs the host, re-issue the first j steps from
 * the untouched input buffers.  Tomography models and models with a likelihood power take K = 1 only. */
int qb_fused_update_multi(const qb_model* model, const qb_expparams* eps, const int64_t* outcomes,
                          int32_t nsteps, uint32_t resample_mask,
                          const double* d_x, int64_t n, const double* d_w_in, double* d_w_out,
                          const double* d_stats_in, double* d_stats_out, double* d_step_stats,
                          const qb_update_ctl* ctl, void* d_ws, size_t ws_bytes, void* stream);

/* Plain likelihood tensor L[o][i][e] (n_o, n, n_e) — Model.likelihood
 * (abstract_model.py:443-468) for the built-in models; `eps`/`outcomes` are HOST arrays. */
int qb_likelihood(const qb_model* model, const qb_expparams* eps, int32_t n_e,
                  const int64_t* outcomes, int32_t n_o,
                  const double* d_x, int64_t n, double* d_L, void* stream);

/* SMCUpdater.hypothetical_update (smc.py:324-386) for every (outcome, experiment) pair, on the device:
 * d_weights[o][e][i] = (w_i L_oie) / norm_oe with the |norm| < eps -> 1 guard of smc.py:369-370, d_norms[o][e] =
 * sum_i w_i L_oie, and optionally d_L[o][e][i] = L_oie (NULL to skip).  w_i is the normalised weight. */
int qb_hypothetical_update(const qb_model* model, const qb_expparams* eps, int32_t n_e,
                           const int64_t* outcomes, int32_t n_o,
                           const double* d_x, const double* d_w, const double* d_stats, int64_t n,
                           double* d_weights, double* d_L, double* d_norms,
                           void* d_ws, size_t ws_bytes, void* stream);

/* Experiment design (SMCUpdater.bayes_risk, smc.py:553-605; expected_information_gain, smc.py:607-657) for ONE
 * experiment `ep` with outcome list `outcomes` (HOST, n_o entries = model.domain(ep).values).  Every outcome's
 * likelihood is evaluated (the reference takes the last one as 1 - sum of the others, smc.py:589, which cancels
 * to +-1e-16 under BinomialModel and then yields NaN gains); a particle of zero likelihood contributes the limit
 * 0 to d_kld instead of 0 * log 0 = NaN.  With h_oi = w_i L_oi:
 *   d_sums[o][0]       = N_o = sum_i h_oi
 *   d_sums[o][1+j]     = sum_i h_oi (x_ij - centre_j)          j < d
 *   d_sums[o][1+d+j]   = sum_i h_oi (x_ij - centre_j)^2
 *   d_kld[o] (if not NULL) = sum_i w_hyp_oi log(w_hyp_oi / w_i),  w_hyp_oi = h_oi / N_o  (smc.py:651)
 * `h_centre` (HOST, d doubles) is any reference point (the posterior mean keeps the variance formula
 * var = C/N - (B/N)^2 free of cancellation).  The (n_o, n) tensors of the reference never exist. */
size_t qb_design_workspace_bytes(int64_t n, int32_t d, int32_t n_o);
int qb_design_sums(const qb_model* model, const qb_expparams* ep, const int64_t* outcomes, int32_t n_o,
                   const double* d_x, const double* d_w, const double* d_stats, int64_t n,
                   const double* h_centre, double* d_sums, double* d_kld,
                   void* d_ws, size_t ws_bytes, void* stream);

/* Model.are_models_valid (test_models.py:109-110, rb.py:149-176,
 * tomography/models.py:143-147): d_valid[i] in {0,1}. */
int qb_are_models_valid(const qb_model* model, const double* d_x, int64_t n, uint8_t* d_valid, void* stream);

/* ---- moments -------------------------------------------------------------- */
size_t qb_moments_workspace_bytes(int64_t n, int32_t d);
/* d_out[0] = sum w, d_out[1..d] = sum w x  (ParticleDistribution.particle_mean,
 * distributions.py:337-348), d_out[1+d .. 1+d+d*d) = sum w x x^T row-major (the
 * einsum of distributions.py:386-387); w = normalised weight.  The host forms
 * cov = E[xx^T] - mu mu^T exactly as distributions.py:388-389. */
int qb_moments(const double* d_x, const double* d_w, const double* d_stats, int64_t n, int32_t d,
               double* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* ---- Liu-West resampler --------------------------------------------------- */
#define QB_SCAN_FAST 0   /* parallel (re-associated) prefix sum                              */
#define QB_SCAN_EXACT 1  /* reproduces np.cumsum's sequential fp64 rounding bit for bit      */
#define QB_SCAN_FAST_GUIDE 2        /* QB_SCAN_FAST + the draw's guide table scattered in the same pass, for
                                       uniforms in [0, 1) (consumed by qb_lw_draw_move / qb_lw_draw_retry)   */
#define QB_SCAN_FAST_GUIDE_SCALED 3 /* same, for draws scaled by the slab's own total (sharded clouds)       */
size_t qb_cdf_workspace_bytes(int64_t n);
/* d_cdf[i] = cumsum of normalised weights (resamplers.py:308). */
int qb_cdf(const double* d_w, const double* d_stats, int64_t n, double* d_cdf, int32_t mode,
           void* d_ws, size_t ws_bytes, void* stream);
/* QB_SCAN_EXACT continued across slabs (SURVEY §8e "Parity mode": the sequential-CDF chain crosses shards): the
 * running sum starts at *d_carry_in (a DEVICE double: the last CDF entry of the slab before this one; 0 for the first
 * slab) instead of at 0, so that the concatenation of the slabs' outputs is np.cumsum of the concatenated weights
 * (resamplers.py:308) bit for bit.  Same workspace as qb_cdf. */
int qb_cdf_chained(const double* d_w, const double* d_stats, int64_t n, double* d_cdf,
                   const double* d_carry_in, void* d_ws, size_t ws_bytes, void* stream);
/* Diagnostics (synchronises): *h_flag = 1 if the last QB_SCAN_EXACT call on this workspace met weights outside
 * the parallel replay's model (negative, NaN, inf) or timed out and re-did the scan with the sequential kernel. */
int qb_cdf_exact_fallback_flag(const void* d_ws, int64_t n, int32_t* h_flag, void* stream);
/* d_js[i] = min(searchsorted(cdf, u[i], side='right'), n-1) (resamplers.py:318-321;
 * the clamp is distributions.py:330-333's — the reference's resampler raises
 * IndexError where the clamp acts).  *d_overflow counts clamped draws.  With a workspace of
 * qb_draw_workspace_bytes(n) the search goes through a power-of-two guide table built on the fly
 * (same indices bit for bit, ~3 instead of ~8 random sectors per draw); d_ws may be NULL. */
size_t qb_draw_workspace_bytes(int64_t n);
int qb_draw(const double* d_cdf, int64_t n, const double* d_u, int64_t n_draw,
            int64_t* d_js, int64_t* d_overflow, void* d_ws, size_t ws_bytes, void* stream);
/* First Liu-West pass (resamplers.py:325-342): for i < n_new
 *   mu_i = a * x_old[js[i]] + (1-a) * mean ;  x_new[i] = mu_i + S @ eps[:, i]
 * `d_eps` is (d, n_new) row-major — the layout of `kernel(n_rvs, k)`.
 * With postselect != 0, d_invalid[i] = !are_models_valid(x_new[i]) and
 * *d_n_invalid is their count.  h_mean (d) and h_S (d*d, row-major, already
 * scaled by h) are HOST arrays. */
/* d_ws: a device buffer of qb_lw_move_workspace_bytes(d) bytes owned by the caller (one per cloud / stream) that
 * receives S and (1-a) mean for the generic-d kernels; qb_lw_move ignores it for d <= 4 (may be NULL there). */
size_t qb_lw_move_workspace_bytes(int32_t d);
int qb_lw_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
               const int64_t* d_js, const double* h_mean, const double* h_S, double a,
               const double* d_eps, int64_t n_new, double* d_x_new,
               int32_t postselect, uint8_t* d_invalid, int64_t* d_n_invalid, void* d_ws, size_t ws_bytes,
               void* stream);
size_t qb_compact_workspace_bytes(int64_t n);
/* Ordered compaction: d_idxs_out = ascending indices i with d_invalid[i] != 0
 * (np.nonzero, resamplers.py:365-367); *d_count = how many. */
int qb_compact_invalid(const uint8_t* d_invalid, int64_t n, int64_t* d_idxs_out, int64_t* d_count,
                       void* d_ws, size_t ws_bytes, void* stream);
/* Retry pass (resamplers.py:327-372) over the k still-invalid particles:
 *   x_new[idxs[r]] = (a * x_old[js[r]] + (1-a) * mean) + S @ eps[:, r],  r < k
 * NOTE js[r], not js[idxs[r]]: the reference re-slices `mus = mus[:k]`
 * (resamplers.py:372), a prefix of the ORIGINAL means; parity replicates it.
 * Sets d_invalid[idxs[r]] to the new validity and *d_n_invalid to the count. */
int qb_lw_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                const int64_t* d_js, const int64_t* d_idxs, int64_t k,
                const double* h_mean, const double* h_S, double a,
                const double* d_eps, double* d_x_new,
                uint8_t* d_invalid, int64_t* d_n_invalid,
                int32_t own_mean /* 0: the reference's js[r] quirk; 1: js[idxs[r]] (sharded clouds) */,
                void* d_ws, size_t ws_bytes, void* stream);

/* Fused first pass for the device-RNG mode, d <= 4: draw + gather + shrink + perturb + validity in ONE launch,
 * without materialising u, js or eps.  New particle i uses uniform element i of Philox stream (seed_u, off_u)
 * and normals eps[m][i] = element m * n_new + i of stream (seed_n, off_n) — exactly the values qb_rng_uniform /
 * qb_rng_normal would have stored, so the result is bit-identical to qb_rng_uniform -> qb_draw -> qb_rng_normal
 * -> qb_lw_move on the same CDF.  `d_cdf` must come from qb_cdf; with use_guide != 0 it must have been built
 * with a QB_SCAN_FAST_GUIDE* mode on the same workspace.  scale_u != 0: draws are u * cdf[n_old-1] (a shard
 * drawing from its own slab).  d_counters[0] = #invalid, d_counters[1] = #clamped draws. */
int qb_lw_draw_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                    const double* d_cdf, const void* d_ws, size_t ws_bytes, int32_t use_guide,
                    const double* h_mean, const double* h_S, double a,
                    uint64_t seed_u, uint64_t off_u, uint64_t seed_n, uint64_t off_n, int32_t scale_u,
                    int64_t n_new, double* d_x_new, int32_t postselect, uint8_t* d_invalid,
                    int64_t* d_counters, void* stream);
/* Retry pass of the fused mode over idxs[0..k): the parent index is re-drawn from uniform element r (the
 * reference's prefix quirk, resamplers.py:372) or idxs[r] (own_mean) of stream (seed_u, off_u); fresh normals
 * eps[m][r] = element m * k + r of stream (seed_n, off_n).  d_counters[0] = #still invalid. */
int qb_lw_draw_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                     const double* d_cdf, const void* d_ws, size_t ws_bytes, int32_t use_guide,
                     const double* h_mean, const double* h_S, double a,
                     uint64_t seed_u, uint64_t off_u, uint64_t seed_n, uint64_t off_n, int32_t scale_u,
                     const int64_t* d_idxs, int64_t k, int32_t own_mean, double* d_x_new,
                     uint8_t* d_invalid, int64_t* d_counters, void* stream);

/* Merge draw for the device-RNG mode, d <= 4: the n_new uniforms are generated already sorted (exponential
 * spacings E_k = -log(1 - u_k) of Philox stream (seed_e, off_e), k = 0 .. n_new, prefix-summed and normalised: exactly
 * the order statistics of n_new i.i.d. uniforms), so the draw is a streaming merge with the CDF instead of n_new
 * random bisections; slot k takes the k-th smallest uniform (the new particles are exchangeable), normals as in
 * qb_lw_draw_move.  Needs n_new <= n_old and a CDF built with a QB_SCAN_FAST_GUIDE* mode on `d_ws` when use_guide
 * != 0 (its tile scratch is reused).  d_parent_inv (n_new int32) receives the parent of every slot found invalid
 * (input of qb_lw_merge_retry, which re-centres on the particle's own parent).  d_u_out / d_js_out (may be NULL):
 * the sorted uniforms and every slot's parent, for tests.  d_counters as in qb_lw_draw_move. */
int qb_lw_merge_move(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                     const double* d_cdf, void* d_ws, size_t ws_bytes, int32_t use_guide,
                     const double* h_mean, const double* h_S, double a,
                     uint64_t seed_e, uint64_t off_e, uint64_t seed_n, uint64_t off_n, int32_t scale_u,
                     int64_t n_new, double* d_x_new, int32_t postselect, uint8_t* d_invalid,
                     int32_t* d_parent_inv, int64_t* d_counters, double* d_u_out, int64_t* d_js_out, void* stream);
int qb_lw_merge_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                      const double* h_mean, const double* h_S, double a, uint64_t seed_n, uint64_t off_n,
                      const int64_t* d_idxs, int64_t k, const int32_t* d_parent_inv, double* d_x_new,
                      uint8_t* d_invalid, int64_t* d_counters, void* stream);

/* Binned multinomial resample for the device-RNG mode, d <= 4 (the default draw of LiuWestResampler(rng='philox',
 * scan='fast')): resamplers.py:266-273 (moments), :308-321 (multinomial draw), :325-372 (shrink, perturb,
 * postselection) and :390-392 (weights 1/n) in three streaming launches, no global CDF, no stored uniforms or
 * indices (csrc/qb_binned.cu describes the factorisation: multinomial counts per bin of 2048 particles, then i.i.d.
 * draws inside each bin from a shared-memory CDF).  The offspring multiset has exactly the reference's law; the
 * slot ORDER differs (grouped by parent bin).
 *   qb_lw_binned_prepare  one pass over (w, x): bin sums -> bin-level CDF, weighted moments into d_moments_out
 *                         (1 + d + d*d doubles, layout of qb_moments; may be NULL) and into the pinned, device-mapped
 *                         host block h_mirror (32 doubles, [31] = tag, written last; may be NULL); then the multinomial
 *                         counts of n_new draws (Philox stream (seed_u, off_u), elements 0..n_new-1), their output
 *                         offsets and the segment list.
 *   qb_lw_binned_move     slot i draws inside its bin with uniform element i of (seed_v, off_v), normals
 *                         eps[m][i] = element m * n_new + i of (seed_n, off_n).  Slots >= split go to
 *                         d_x_new2[(i - split)] when d_x_new2 != NULL (a shard's surplus rows).  d_w_new != NULL:
 *                         also stores the new weights 1/n_global and their stats block d_stats_new.  Invalid slots
 *                         are appended to d_list (n_new int64: slot | parent << 32); h_mirror (8 doubles, 32-byte
 *                         aligned, device-accessible: pinned host or device memory) receives {#invalid, #clamped,
 *                         #drawn, tag}.  retry_rounds > 0 (with postselect) queues qb_lw_binned_retry right behind it
 *                         (rounds = retry_rounds, stream offset off_n + round_stride), reporting into h_mirror[4..8).
 *                         d_js_out (may be NULL): the parent of every slot (tests).
 *   qb_lw_binned_retry    up to `rounds` fresh perturbations per still-invalid list entry in ONE launch (round j: normals
 *                         element m * n_new + slot of (seed_n, off_n + j * round_stride)), stopping at the first valid
 *                         one; resolved entries become -1.  Re-centred on the slot's own parent (own_mean != 0), or —
 *                         the reference's law: resamplers.py:372 re-slices `mus = mus[:k]`, so the r-th still-invalid
 *                         particle is re-centred on the r-th ORIGINAL draw, an i.i.d. draw from the weighted cloud —
 *                         on the parent of a uniformly random slot (d_parents: the n_new int32 parent indices the move
 *                         stored; uniform element `slot` of stream (seed_v, off_n + j * round_stride)), fresh each round.  The list length is
 *                         read on the device (a counter in the workspace), so the call may be queued right behind the
 *                         move.  h_mirror (4 doubles) receives {#still invalid, most rounds used, list length, tag}.
 * Workspace: qb_lw_binned_workspace_bytes(n_old, n_new), ZERO-INITIALISED once, private to these three calls. */
size_t qb_lw_binned_workspace_bytes(int64_t n_old, int64_t n_new);
/* qb_lw_binned_prepare = qb_lw_binned_sums (pass 1) + qb_lw_binned_count (pass 2).  A sharded cloud calls them
 * separately: the shard masses of pass 1 decide how many offspring n_new this slab produces. */
int qb_lw_binned_sums(const double* d_x, const double* d_w, const double* d_stats, int64_t n_old, int32_t d,
                      double* d_moments_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes, void* stream);
/* How pass 2 obtains the multinomial counts: QB_COUNT_HISTOGRAM locates n_new Philox uniforms (elements 0..n_new-1 of
 * (seed_u, off_u)) in the bin-level CDF held in shared memory; QB_COUNT_TREE splits n_new down a binary tree over the
 * bins with one exact Binomial variate per node (inversion / BTPE, Philox counters off_u + 32 k of node k) — <= 2T
 * variates whatever n_new.  Same law (Multinomial(n_new; bin masses)); both are pure functions of (weights, seed,
 * offset).  QB_COUNT_AUTO (default of the Python host): the histogram while its tables fit in shared memory
 * (n_old <= 3.4e7), the tree beyond. */
#define QB_COUNT_AUTO 0
#define QB_COUNT_HISTOGRAM 1
#define QB_COUNT_TREE 2
/* Sharded cloud, between pass 1 and pass 3: `d_rows` holds n_ranks rows of 1 + d + d*d moment sums (the
 * all-gathered `d_moments_out` of every rank's qb_lw_binned_sums, globally normalised weights).  One thread sums them
 * in rank order — bit-identical on every rank — and derives the Liu-West constants exactly like
 * qb_lw_binned_resample's first kernel (same workspace slot: a following qb_lw_binned_move / _retry with h_mean = h_S =
 * NULL uses them).  h_mirror (may be NULL): [0 ..) the global moments, [29] covariance flag, [30] sqrtm error,
 * [31] tag, as qb_lw_binned_resample publishes them. */
int qb_lw_binned_shard_consts(const double* d_rows, int32_t n_ranks, int32_t d, double a, double h,
                              double zero_cov_comp, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                              void* stream);
int qb_lw_binned_count(int64_t n_old, int64_t n_new, uint64_t seed_u, uint64_t off_u, int32_t mode, void* d_ws,
                       size_t ws_bytes, void* stream);
/* The tree's sampler on its own (tests): d_out[i] ~ Binomial(n, p), i < count, from counters off + 32 i of stream seed. */
int qb_binomial_sample(int64_t n, double p, int64_t count, uint64_t seed, uint64_t off, int64_t* d_out, void* stream);
int qb_lw_binned_prepare(const double* d_x, const double* d_w, const double* d_stats, int64_t n_old, int32_t d,
                         int64_t n_new, uint64_t seed_u, uint64_t off_u, int32_t count_mode, double* d_moments_out,
                         double* h_mirror, double tag, void* d_ws, size_t ws_bytes, void* stream);
int qb_lw_binned_move(const qb_model* model, const double* d_x_old, const double* d_w, const double* d_stats,
                      int64_t n_old, int32_t d, const double* h_mean, const double* h_S, double a,
                      uint64_t seed_v, uint64_t off_v, uint64_t seed_n, uint64_t off_n, int64_t n_new,
                      double* d_x_new, int64_t split, double* d_x_new2, double* d_w_new, int64_t n_global,
                      double* d_stats_new, int32_t postselect, int32_t retry_rounds, int32_t own_mean, int64_t* d_list,
                      int32_t* d_parents, int64_t* d_js_out, double* h_mirror, double tag, void* d_ws, size_t ws_bytes,
                      void* stream);
int qb_lw_binned_retry(const qb_model* model, const double* d_x_old, int64_t n_old, int32_t d,
                       const double* h_mean, const double* h_S, double a, uint64_t seed_n, uint64_t off_n,
                       uint64_t round_stride, int32_t rounds, int32_t own_mean, uint64_t seed_v, int64_t n_new,
                       double* d_x_new, int64_t split, double* d_x_new2, int64_t* d_list, const int32_t* d_parents,
                       double* h_mirror, double tag, void* d_ws, size_t ws_bytes, void* stream);

/* The whole resample queued by ONE call, with no host decision in between: pass 1 additionally derives the Liu-West
 * constants on the device — cov = E[xx^T] - mu mu^T (distributions.py:388-389), a zero Frobenius norm replaced by
 * zero_cov_comp * I (resamplers.py:288-293), S = h * sqrtm_psd(cov) (utils.py:593-607: symmetric eigendecomposition,
 * here cyclic Jacobi; eigenvalues <= 0 clipped), (1 - a) * mean — and pass 3 reads them from the workspace.
 * h_mirror: 40 doubles, 32-byte aligned, device-accessible: [0, 1+d+d*d) the moments, [29] covariance flag (0 fine,
 * 1 zero norm replaced, 2 not finite), [30] || S0 S0 - cov ||_F, [31] = tag once pass 1 has published; [32, 36) the
 * move's {#invalid, #clamped, #drawn, tag}; [36, 40) the queued retry's {#still invalid, rounds used, list length, tag}.
 * The host checks the flags and raises the reference's warnings / ResamplerError after the fact. */
int qb_lw_binned_resample(const qb_model* model, const double* d_x, const double* d_w, const double* d_stats,
                          int64_t n_old, int32_t d, int64_t n_new, double a, double h, double zero_cov_comp,
                          uint64_t seed, uint64_t off_u, int32_t count_mode, uint64_t off_v, uint64_t seed_n,
                          uint64_t off_n, double* d_x_new, double* d_w_new, int64_t n_global, double* d_stats_new,
                          int32_t postselect, int32_t retry_rounds, int32_t own_mean, int64_t* d_list,
                          int32_t* d_parents, double* d_moments_out, double* h_mirror, double tag, void* d_ws,
                          size_t ws_bytes, void* stream);

/* ---- sharded cloud: peer mailboxes and the resample exchange ------------------- */
#define QB_IPC_HANDLE_BYTES 64
/* cudaMalloc + zero a mailbox of 2 * n_ranks * QB_MAILBOX_ROW doubles (the only entry points that allocate). */
int qb_mailbox_create(int32_t n_ranks, double** d_mailbox);
int qb_mailbox_destroy(double* d_mailbox);
/* CUDA IPC plumbing so that a peer PROCESS can map the mailbox (handles travel over torch.distributed). */
int qb_ipc_get_handle(const void* d_ptr, unsigned char handle[QB_IPC_HANDLE_BYTES]);
int qb_ipc_open_handle(const unsigned char handle[QB_IPC_HANDLE_BYTES], void** d_ptr);
int qb_ipc_close_handle(void* d_ptr);
/* Resample routing.  bounds[r] = global CDF value at the START of shard r (bounds[n_ranks] = total), HOST
 * array.  For each of the n uniforms: owner = the shard whose CDF range contains it.
 *   d_counts[r]  (int64, zeroed by the call) = how many of my draws shard r owns
 * then, given the exclusive prefix d_bucket_start of those counts (HOST int64[n_ranks]):
 *   d_req[pos]   = u - bounds[owner]   (the owner-local CDF coordinate, bucketed by owner)
 *   d_perm[i]    = pos                 (where draw i's row will come back) */
int qb_shard_classify(const double* d_u, int64_t n, const double* h_bounds, int32_t n_ranks,
                      int32_t* d_owner, int64_t* d_counts, void* stream);
int qb_shard_bucket(const double* d_u, const int32_t* d_owner, int64_t n, const double* h_bounds,
                    const int64_t* h_bucket_start, int32_t n_ranks, int64_t* d_cursor /* n_ranks, scratch */,
                    double* d_req, int64_t* d_perm, void* stream);
/* d_out[i][:] = d_x[d_js[i]][:]  — the owner's reply to a batch of resolved requests. */
int qb_gather_rows(const double* d_x, int32_t d, const int64_t* d_js, int64_t n, double* d_out, void* stream);

/* ---- tomography canonicalize ---------------------------------------------- */
/* TomographyModel.canonicalize (tomography/models.py:149-209): per particle
 * rho = sum_a x_a conj(B_a); eigendecompose (Hermitian Jacobi); if any
 * eigenvalue < 0 clip, rebuild and project back x_a = Re sum_ij B_a[ij] rho'[ij];
 * then divide by x_0 sqrt(dim) unless allow_subnormalized.  `d_basis` is the
 * (dim^2, dim, dim) complex128 basis tensor (interleaved re/im), dim in {2,3,4}. */
int qb_tomo_canonicalize(double* d_x, int64_t n, int32_t dim, const double* d_basis,
                         int32_t allow_subnormalized, void* stream);
/* Same result, for large clouds: a screening pass certifies the comfortably positive-definite particles with an
 * LDL^H factorisation (for them canonicalize is the identity + the renormalising division) and flags the rest;
 * the flagged ones are compacted (qb_compact_invalid) and only they go through the eigendecomposition.
 * Scratch: d_flags (n bytes), d_idxs (n int64), d_count (one int64, stays on the device) and a workspace of
 * qb_compact_workspace_bytes(n). */
int qb_tomo_canonicalize_screened(double* d_x, int64_t n, int32_t dim, const double* d_basis,
                                  int32_t allow_subnormalized, uint8_t* d_flags, int64_t* d_idxs,
                                  int64_t* d_count, void* d_ws, size_t ws_bytes, void* stream);

/* Same two calls for particles stored with a row pitch of `ld` >= dim^2 doubles (DiffusiveTomographyModel keeps its
 * diffusion rate behind the dim^2 state parameters, tomography/models.py:251-255: only the first dim^2 entries of a
 * row are canonicalised, the rest passes through). */
int qb_tomo_canonicalize_ld(double* d_x, int64_t n, int32_t dim, int32_t ld, const double* d_basis,
                            int32_t allow_subnormalized, void* stream);
int qb_tomo_canonicalize_screened_ld(double* d_x, int64_t n, int32_t dim, int32_t ld, const double* d_basis,
                                     int32_t allow_subnormalized, uint8_t* d_flags, int64_t* d_idxs,
                                     int64_t* d_count, void* d_ws, size_t ws_bytes, void* stream);

/* ---- read-side estimators ------------------------------------------------------------------------ */
/* ParticleDistribution.est_entropy (distributions.py:457-465): *d_out = -sum over w_i > 0 of w_i log w_i (w normalised).
 * Workspace: qb_readside_workspace_bytes(). */
size_t qb_readside_workspace_bytes(void);
int qb_weights_entropy(const double* d_w, const double* d_stats, int64_t n, double* d_out, void* d_ws, size_t ws_bytes,
                       void* stream);
/* est_credible_region (distributions.py:558-614) by radix SELECTION instead of a sort: one pass histograms the
 * normalised weights whose bit pattern above (shift + nbits) equals `prefix` by their nbits-wide digit at `shift`
 * (nbits <= 11): d_mass[b] = their weight mass, d_count[b] = how many (2048 entries each, zeroed by the call).
 * Non-negative doubles order like their patterns; weights <= 0 or NaN are skipped (they carry no mass). */
int qb_weight_mass_hist(const double* d_w, const double* d_stats, int64_t n, int32_t shift, int32_t nbits,
                        uint64_t prefix, double* d_mass, uint64_t* d_count, void* stream);
/* d_flags[i] = 1 iff the pattern of the normalised weight i is > tau_bits (mode 0) or == tau_bits (mode 1); input of
 * qb_compact_invalid, whose index list + qb_gather_rows bring only the region's members to the host. */
int qb_weights_select(const double* d_w, const double* d_stats, int64_t n, uint64_t tau_bits, int32_t mode,
                      uint8_t* d_flags, void* stream);

/* ---- small clouds, parity mode: the first Liu-West pass in ONE single-CTA launch --------------------------------- */
/* For n_old <= QB_SMALL_MAX, d <= 4 (the reference's own CPU-sized runs, BASELINE config C1, are launch-latency bound
 * on a GPU).  One CTA: moments -> covariance / zero-norm replacement / matrix square root (as qb_lw_binned_resample's
 * pass 1) -> np.cumsum of the normalised weights, sequential fp64 (resamplers.py:308) -> d_js[i] =
 * min(searchsorted(cdf, d_u[i], 'right'), n_old - 1) (resamplers.py:318-321) -> x_new[i] = a x[js[i]] + (1 - a) mean +
 * S eps[:, i] (resamplers.py:325-332; d_eps is (d, n_new) row-major) -> d_invalid flags, d_counters[0] = their count,
 * d_counters[1] = clamped draws -> optionally the uniform weights 1 / n_new and their stats block.  d_u / d_eps are the
 * HOST-drawn legacy variates (np.random), uploaded by the caller.  The retry iterations use qb_compact_invalid +
 * qb_lw_retry on the same buffers.  h_mirror (pinned, 64 doubles, 32-byte aligned): [0..) moments, [29] covariance
 * flag, [30] sqrtm error, [31] tag | [32] invalid, [33] clamped, [34] n_new, [35] tag | [40..56) S (times h),
 * [56..60) (1 - a) mean. */
#define QB_SMALL_MAX 4096
int qb_lw_small_resample(const qb_model* model, const double* d_x, const double* d_w, const double* d_stats,
                         int64_t n_old, int32_t d, double a, double h, double zero_cov_comp,
                         const double* d_u, const double* d_eps, int64_t n_new, double* d_x_new,
                         int64_t* d_js, uint8_t* d_invalid, int64_t* d_counters, double* d_w_new /* may be NULL */,
                         double* d_stats_new /* may be NULL */, int32_t postselect, double* d_moments_out /* may be NULL */,
                         double* h_mirror, double tag, void* stream);

/* ---- time-dependent and noisy decorators ------------------------------------------------------- */
/* Model.update_timestep for the random-walk decorators, in place after an update has been committed (smc.py:447-449).
 * For each particle i and each of the n_rw walking parameters c (HOST arrays of n_rw entries):
 *   QB_WALK_ADD      x[i][idx[c]] += (mult * z[i][zcol[c]])                        z holds the steps themselves
 *                    (RandomWalkModel, derived_models.py:733-741, steps drawn by the model's own distribution)
 *   QB_WALK_FIXED    x[i][idx[c]] += (mult * (scale[c] * z[i][zcol[c]]))           fixed diagonal covariance
 *                    (GaussianRandomWalkModel, derived_models.py:925-929,944-962)
 *   QB_WALK_LEARNED  x[i][idx[c]] += (mult * ((x[i][sidx[c]] * pre) * z[i][zcol[c]]))   per-particle scale: learned
 *                    sigma (derived_models.py:926, pre = 1) or DiffusiveTomographyModel's eps * sqrt(t)
 *                    (tomography/models.py:261-266, pre = sqrt(t), mult = 1)
 * with one rounding per reference ufunc.  `d_z` is (n, kz) row-major standard normals (or steps). */
#define QB_WALK_ADD 0
#define QB_WALK_FIXED 1
#define QB_WALK_LEARNED 2
int qb_walk_step(double* d_x, int64_t n, int32_t d, int32_t n_rw, const int32_t* h_idx, const int32_t* h_zcol,
                 int32_t mode, const double* h_scale, const int32_t* h_sidx, double pre, double mult,
                 const double* d_z, int32_t kz, void* stream);
/* PoisonedModel.likelihood's noise (derived_models.py:188-204) on a likelihood vector, in place:
 *   L[i] = clip(L[i] + z[i] * sigma_i, 0, 1),  sigma_i = tol (mode 0, ALE) or
 *   sqrt(L[i] * (1 - L[i]) / denom) (mode 1, MLE: binom_est_error, utils.py:683-688, denom = N + 2 hedge + 1). */
int qb_poison_likelihood(double* d_L, int64_t n, const double* d_z, int32_t mode, double tol, double denom,
                         void* stream);

/* ---- device RNG (throughput mode; counter-based Philox4x32-10) ------------- */
/* d_out[i] = uniform [0,1) with 53 random bits, element i of stream (seed, offset). */
int qb_rng_uniform(double* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream);
/* d_out[i] = standard normal (Box-Muller on two 53-bit uniforms). */
int qb_rng_normal(double* d_out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* ---- device RNG (parity mode at scale; NumPy's legacy MT19937 stream) ------ */
/* Continue the global np.random stream the reference draws from
 * (np.random.random, resamplers.py:319; np.random.randn, resamplers.py:332)
 * on the device.  `h_key`/`pos` (and `has_gauss`/`cached`) are the fields of
 * np.random.get_state(); the *_out values are what np.random.set_state() takes
 * afterwards.  Uniforms are bit-identical to NumPy's; for normals the accepted
 * set, the words consumed and the final state are exact and the values are
 * within 1 ulp (device log()).  Both calls synchronise the stream (they return
 * host state).  Workspace for n uniforms and/or m normals: */
size_t qb_mt19937_workspace_bytes(int64_t n_uniform, int64_t n_normal);
int qb_mt19937_uniform(const uint32_t* h_key, int32_t pos, int64_t n, double* d_out, uint32_t* h_key_out,
                       int32_t* pos_out, void* d_ws, size_t ws_bytes, void* stream);
int qb_mt19937_normal(const uint32_t* h_key, int32_t pos, int32_t has_gauss, double cached, int64_t m,
                      double* d_out, uint32_t* h_key_out, int32_t* pos_out, int32_t* has_gauss_out,
                      double* cached_out, void* d_ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QINFER_B200_H */
