// Per-particle likelihoods and validity tests of the built-in models.
//
// Compiled with --fmad=false: the reference evaluates these expressions with
// one rounding per NumPy ufunc, so every multiply/add here must round
// separately (no silent FMA contraction); explicit fma() is used only where
// the reference's BLAS call does.
#pragma once
#include "qb_common.cuh"
#include "qb_trig.cuh"

namespace qb {

// Device-side view of one (experiment, outcome) pair, prepared on the host.
struct ExpView {
    double t, w_;      // precession
    double m;          // RB: sequence length as float64 (NumPy casts uint -> float64 for p ** m, rb.py:193)
    int32_t reference; // interleaved RB
    int32_t outcome0;  // 1 if the two-outcome label == 0 (abstract_model.py:683-686: anything else is "1")
    double n_meas, k;  // BinomialModel: shots and observed count
    double logc;       // lgamma(n+1) - lgamma(k+1) - lgamma(n-k+1), hoisted per update
    int32_t k_in_range; // 0 <= k <= n
    int32_t fast_binom; // fast_math: n_meas small enough for C(n,k) p^k (1-p)^(n-k) by integer powers
    double binc;       // fast_math: the binomial coefficient C(n, k) (exact below 2^53)
    uint32_t m_int, k_int, nk_int, pad;   // fast_math: the integer exponents
};

struct ModelView {
    int32_t kind, d, binomial, interleaved;
    double min_freq;
    double like_pow;   // MLEModel: L ** like_pow (1 = plain)
    int32_t d_extra, extra_rule;  // trailing parameters the likelihood ignores, and their validity rule
    int32_t fast_math, pad0;      // integer powers instead of pow / log / exp (qb_model.fast_math)
};

__host__ inline ExpView make_exp_view(const qb_model& m, const qb_expparams& ep, int64_t outcome) {
    ExpView v;
    v.t = ep.t;
    v.w_ = ep.w_;
    v.m = static_cast<double>(static_cast<uint64_t>(ep.m));
    v.reference = ep.reference;
    v.outcome0 = (outcome == 0) ? 1 : 0;
    v.n_meas = static_cast<double>(ep.n_meas);
    v.k = static_cast<double>(outcome);
    v.k_in_range = (outcome >= 0 && outcome <= ep.n_meas) ? 1 : 0;
    v.logc = 0.0;
    if (m.binomial && v.k_in_range) {
        v.logc = lgamma(v.n_meas + 1.0) - lgamma(v.k + 1.0) - lgamma(v.n_meas - v.k + 1.0);
    }
    v.pad = 0;
    v.fast_binom = 0;
    v.binc = 1.0;
    v.m_int = static_cast<uint32_t>(static_cast<uint64_t>(ep.m));
    v.k_int = v.nk_int = 0;
    if (m.fast_math && m.binomial && v.k_in_range && ep.n_meas <= 56) {
        // C(n, k) by the multiplicative formula in extended precision: exact integers below 2^53 for n <= 56
        const int64_t k = (outcome < ep.n_meas - outcome) ? outcome : ep.n_meas - outcome;
        long double c = 1.0L;
        for (int64_t i = 1; i <= k; ++i) c = c * static_cast<long double>(ep.n_meas - k + i) / static_cast<long double>(i);
        v.binc = static_cast<double>(roundl(c));
        v.k_int = static_cast<uint32_t>(outcome);
        v.nk_int = static_cast<uint32_t>(ep.n_meas - outcome);
        v.fast_binom = 1;
    }
    return v;
}

// x ** e for an unsigned integer exponent by squaring: <= 2 log2(e) multiplications, each rounding once.  The exponent is
// uniform over the launch, so the loop does not diverge.
__device__ __forceinline__ double ipow(double x, uint32_t e) {
    double r = 1.0;
    while (e) {
        if (e & 1u) r *= x;
        e >>= 1;
        if (e) x *= x;
    }
    return r;
}

__host__ inline ModelView make_model_view(const qb_model& m) {
    ModelView v;
    v.kind = m.kind;
    v.d = m.d;
    v.binomial = m.binomial;
    v.interleaved = m.interleaved;
    v.min_freq = m.min_freq;
    v.like_pow = (m.likelihood_power == 0.0) ? 1.0 : m.likelihood_power;
    v.d_extra = m.d_extra;
    v.extra_rule = m.extra_rule;
    v.fast_math = m.fast_math;
    v.pad0 = 0;
    return v;
}

// pr0 of the underlying two-outcome model for one particle.  `row(c)` returns
// model parameter c of this particle, `meas(c)` the tomography coefficient c.
// `rot` rotates the summation start so that lanes of a warp touch different
// shared-memory banks when rows are 128-B apart (d = 16, 64).
template <int KIND, typename Row, typename Meas>
__device__ __forceinline__ double model_pr0(const ModelView& mv, const ExpView& ev, Row row, Meas meas, int rot) {
    if (KIND == QB_MODEL_PRECESSION) {
        // test_models.py:134-140: cos(t * (omega - w_) / 2) ** 2
        double dw = row(0) - ev.w_;
        double th = (ev.t * dw) / 2.0;
        if (fabs(th) < TRIG_FAST_LIMIT) return sincos_squared_fast(th).cos2;
        double c = cos(th);  // huge or non-finite argument: library path (Payne-Hanek)
        return c * c;
    } else if (KIND == QB_MODEL_RB) {
        // rb.py:178-195: 1 - (A * p**m + B); interleaved: p <- p or p~ * p
        double p, A, B;
        if (mv.interleaved) {
            double pt = row(0);
            p = row(1);
            A = row(2);
            B = row(3);
            if (!ev.reference) p = pt * p;
        } else {
            p = row(0);
            A = row(1);
            B = row(2);
        }
        double pm = mv.fast_math ? ipow(p, ev.m_int) : pow(p, ev.m);
        return 1.0 - (A * pm + B);
    } else if (KIND == QB_MODEL_COIN) {
        return row(0);  // test_models.py:323: pr0 is the coin's bias itself
    } else {
        // tomography/models.py:214-226: pr1 = clip(<meas, x>, 0, 1); pr0 = 1 - pr1
        const int d = mv.d;
        double acc = 0.0;
        int c = rot % d;
        for (int j = 0; j < d; ++j) {
            acc = fma(meas(c), row(c), acc);
            c = (c + 1 == d) ? 0 : c + 1;
        }
        acc = fmin(fmax(acc, 0.0), 1.0);
        return 1.0 - acc;
    }
}

// Binomial pmf, SciPy's closed form (derived_models.py:324-327 -> utils.py:106-111):
// exp(logC + xlogy(k, p) + xlog1py(n - k, -p)).
__device__ __forceinline__ double binom_pmf(const ExpView& ev, double p) {
    if (!ev.k_in_range) return 0.0;
    if (ev.fast_binom) return (ev.binc * ipow(p, ev.k_int)) * ipow(1.0 - p, ev.nk_int);
    double t1 = (ev.k == 0.0) ? 0.0 : ev.k * log(p);
    double nk = ev.n_meas - ev.k;
    double t2 = (nk == 0.0) ? 0.0 : nk * log1p(-p);
    return exp(ev.logc + t1 + t2);
}

// (pr0, pr1) of the underlying two-outcome model.  pr1 = 1 - pr0 as in
// abstract_model.py:665-686, except for the precession model where sin^2 is available directly
// (it equals 1 - cos^2 to within the reference's own rounding and is relatively accurate near 0).
template <int KIND, typename Row, typename Meas>
__device__ __forceinline__ void model_pr01(const ModelView& mv, const ExpView& ev, Row row, Meas meas, int rot,
                                           double& pr0, double& pr1) {
    if (KIND == QB_MODEL_PRECESSION) {
        const double dw = row(0) - ev.w_;
        const double th = (ev.t * dw) / 2.0;
        if (fabs(th) < TRIG_FAST_LIMIT) {
            const SinCosSq sc = sincos_squared_fast(th);
            pr0 = sc.cos2;
            pr1 = sc.sin2;
            return;
        }
        const double c = cos(th);
        pr0 = c * c;
        pr1 = 1.0 - pr0;
        return;
    }
    pr0 = model_pr0<KIND>(mv, ev, row, meas, rot);
    pr1 = 1.0 - pr0;
}

template <int KIND, bool BINOM, typename Row, typename Meas>
__device__ __forceinline__ double model_likelihood_plain(const ModelView& mv, const ExpView& ev, Row row, Meas meas,
                                                         int rot) {
    if (KIND == QB_MODEL_PRECESSION && !BINOM) {
        // two-outcome precession: one select between sin^2(r) and 1 - sin^2(r)
        const double dw = row(0) - ev.w_;
        const double th = (ev.t * dw) / 2.0;
        if (fabs(th) < TRIG_FAST_LIMIT) return cos2_or_sin2_fast(th, ev.outcome0);
        const double c = cos(th);  // huge or non-finite argument: library path (Payne-Hanek)
        const double pr0 = c * c;
        return ev.outcome0 ? pr0 : 1.0 - pr0;
    }
    double pr0, pr1;
    model_pr01<KIND>(mv, ev, row, meas, rot, pr0, pr1);
    // underlying.likelihood([1], ...) = 1 - pr0 exactly as derived_models.py:318-321 forms it: log(pr1)
    // amplifies the last bit of a small pr1, so the binomial path keeps the reference's rounding sequence
    if (BINOM) return binom_pmf(ev, 1.0 - pr0);
    return ev.outcome0 ? pr0 : pr1;
}

// The likelihood the updater sees: the (binomial) model's, raised to MLEModel's power when one is set
// (derived_models.py:701-703: `L ** self._pow`, one pow() per element like NumPy's).  pow() stays out of line so
// that the plain models' inner loops keep their registers and instruction footprint.
static __device__ __noinline__ double likelihood_pow(double L, double g) { return pow(L, g); }

template <int KIND, bool BINOM, typename Row, typename Meas>
__device__ __forceinline__ double model_likelihood(const ModelView& mv, const ExpView& ev, Row row, Meas meas,
                                                   int rot) {
    const double L = model_likelihood_plain<KIND, BINOM>(mv, ev, row, meas, rot);
    return (mv.like_pow == 1.0) ? L : likelihood_pow(L, mv.like_pow);
}

// Model.are_models_valid for one particle.
template <typename Row>
__device__ __forceinline__ bool model_valid_base(const ModelView& mv, Row row);

// ... including the trailing parameters of a decorator (learned random-walk scales, diffusion rate)
template <typename Row>
__device__ __forceinline__ bool model_valid(const ModelView& mv, Row row) {
    bool ok = model_valid_base(mv, row);
    if (mv.extra_rule == 1) {          // derived_models.py:883-892: every learned step scale >= 0
        for (int c = mv.d - mv.d_extra; c < mv.d; ++c) ok = ok && (row(c) >= 0.0);
    } else if (mv.extra_rule == 2) {   // tomography/models.py:245-249: diffusion rate > 0
        ok = ok && (row(mv.d - 1) > 0.0);
    }
    return ok;
}

template <typename Row>
__device__ __forceinline__ bool model_valid_base(const ModelView& mv, Row row) {
    if (mv.kind == QB_MODEL_PRECESSION) {
        return row(0) > mv.min_freq;  // test_models.py:109-110
    } else if (mv.kind == QB_MODEL_RB) {
        // rb.py:149-176 (the interleaved branch names the first parameter p_C and tests it as such)
        if (mv.interleaved) {
            double pc = row(0), p = row(1), A = row(2), B = row(3);
            return (0.0 <= p) && (p <= 1.0) && (0.0 <= pc) && (pc <= 1.0) && (0.0 <= A) && (A <= 1.0) &&
                   (0.0 <= B) && (B <= 1.0) && (A + B <= 1.0) && (A * p + B <= 1.0) && (A * pc + B <= 1.0);
        }
        double p = row(0), A = row(1), B = row(2);
        return (0.0 <= p) && (p <= 1.0) && (0.0 <= A) && (A <= 1.0) && (0.0 <= B) && (B <= 1.0) &&
               (A + B <= 1.0) && (A * p + B <= 1.0);
    }
    if (mv.kind == QB_MODEL_COIN) return (row(0) >= 0.0) && (row(0) <= 1.0);  // test_models.py:303-304
    return true;  // tomography/models.py:143-147
}

}  // namespace qb
