"""CPU oracle for the SMC Bayes-update + Liu-West resample hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module, and there only as the checker or
as the timed CPU baseline.  The product (``python-qinfer_b200/``) never routes
through it and has no CPU fallback.

What this is: a NumPy restatement of the reference's algorithm
(QInfer/python-qinfer @ 8170c84, paths relative to /root/reference/src/qinfer)
for exactly the functions SURVEY.md §8(a) lists, plus the "next" rows built so far (§8 f2: bayes_risk /
expected_information_gain with CoinModel; f4: MLEModel, PoisonedModel, RandomWalkModel, GaussianRandomWalkModel).  Every function cites the
reference lines it follows and performs the same floating-point operations in
the same order on the same array shapes, so that under an identical legacy
``np.random`` seed it reproduces the reference bit for bit.

Pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported
in the build container through the shims in ``tests/golden/ref_import.py``) and
this oracle side by side on the same seeds and commits the reference's outputs
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle
against those vectors bit-exactly (weights, resample indices, locations,
moments) on every CPU test run.  The reference itself holds no golden vectors
for this path; its behavioural tests (tests/test_smc.py:109-137,
tests/test_precession_model.py:86-110, tests/test_distributions.py:652-706,
tests/test_utils.py:132-152, tests/test_metrics.py:40-120) are restated in ``tests/test_gpu_updater.py``,
``tests/test_host_logic.py`` and ``tests/test_oracle_golden.py``.

Third-party arithmetic the reference delegates (not under /root/reference):
  * ``scipy.stats.binom(n, p).pmf(k)`` (utils.py:106-111), SciPy un-pinned by
    the reference (requirements.txt:2); this image has SciPy 1.18.1 whose pmf
    is Boost.Math ``ibeta_derivative``.  The oracle calls the same SciPy entry
    point; ``binomial_pmf_restated`` is the published closed form
    exp(lgamma(n+1)-lgamma(k+1)-lgamma(n-k+1)+k log p+(n-k) log1p(-p)).
  * ``scipy.linalg.eigh`` (utils.py:598) and ``np.linalg.eig``
    (tomography/models.py:184): called as the reference calls them.
  * NumPy legacy MT19937 ``np.random.random`` / ``randn``, ``np.cumsum``
    (sequential), ``ndarray.searchsorted`` (resamplers.py:308-332).
"""
from __future__ import division

import itertools
import warnings
from functools import reduce

import numpy as np
import scipy.linalg
import scipy.stats

EPS = np.spacing(1)


class ApproximationWarning(RuntimeWarning):
    """Same role as qinfer._exceptions.ApproximationWarning (_exceptions.py:54-63)."""


class ResamplerWarning(RuntimeWarning):
    """Same role as qinfer._exceptions.ResamplerWarning (_exceptions.py:65-70)."""


class ResamplerError(RuntimeError):
    """Same role as qinfer._exceptions.ResamplerError (_exceptions.py:72-76)."""

    def __init__(self, msg, cause=None):
        super(ResamplerError, self).__init__(msg)
        self._cause = cause


# ---------------------------------------------------------------------------
# Likelihoods
# ---------------------------------------------------------------------------

def two_outcome_likelihood(outcomes, pr0):
    """abstract_model.py:665-686 — row i is pr0 if outcomes[i]==0 else 1-pr0.

    ``pr0`` has shape (n_models, n_expparams); result (n_outcomes, n_models,
    n_expparams).  Any outcome label other than 0 selects ``1 - pr0``.
    """
    outcomes = np.atleast_1d(np.asarray(outcomes))
    p0 = pr0[np.newaxis, ...]
    p1 = 1 - p0
    return np.concatenate([p0 if outcomes[i] == 0 else p1 for i in range(outcomes.shape[0])])


def precession_pr0(modelparams, t, w_=0.0):
    """test_models.py:134-140 — pr0 = cos(t (omega - w_) / 2) ** 2."""
    if modelparams.ndim == 1:
        modelparams = modelparams[:, np.newaxis]
    t = np.atleast_1d(np.asarray(t, dtype=float))
    dw = modelparams - w_
    pr0 = np.zeros((modelparams.shape[0], t.shape[0]))
    pr0[:, :] = np.cos(t * dw / 2) ** 2
    return pr0


def rb_pr0(modelparams, m, interleaved=False, reference=None):
    """rb.py:178-195 — pr0 = 1 - (A p**m + B); interleaved picks p or p~ p."""
    m = np.atleast_1d(np.asarray(m))
    if interleaved:
        p_tilde, p, A, B = modelparams.T[:, :, np.newaxis]
        p_C = p_tilde * p
        p = np.where(np.atleast_1d(reference)[np.newaxis, :], p, p_C)
    else:
        p, A, B = modelparams.T[:, :, np.newaxis]
    mm = m[np.newaxis, :]
    pr0 = np.zeros((modelparams.shape[0], m.shape[0]))
    pr0[:, :] = 1 - (A * (p ** mm) + B)
    return pr0


def rb_valid(modelparams, interleaved=False):
    """rb.py:149-176 — the box/affine inequalities on (p, A, B) [and p~]."""
    if interleaved:
        p_C, p, A, B = modelparams.T
        return np.all([0 <= p, p <= 1, 0 <= p_C, p_C <= 1, 0 <= A, A <= 1, 0 <= B, B <= 1,
                       A + B <= 1, A * p + B <= 1, A * p_C + B <= 1], axis=0)
    p, A, B = modelparams.T
    return np.all([0 <= p, p <= 1, 0 <= A, A <= 1, 0 <= B, B <= 1, A + B <= 1, A * p + B <= 1], axis=0)


def binomial_pmf(n, k, p):
    """utils.py:106-111 — scipy.stats.binom(n, p).pmf(k), the same call."""
    return scipy.stats.binom(n, p).pmf(k)


def binomial_pmf_restated(n, k, p):
    """Closed form of the same pmf (SciPy's own ``_logpmf`` formula); used to
    cross-check the third-party call, agreement ~1e-13 relative."""
    from scipy.special import gammaln, xlogy, xlog1py
    n = np.asarray(n, dtype=float)
    k = np.asarray(k, dtype=float)
    logc = gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1)
    return np.exp(logc + xlogy(k, p) + xlog1py(n - k, -p))


def tomography_pr1(modelparams, meas):
    """tomography/models.py:214-224 — pr1 = clip(sum_i meas[e,i] x[m,i], 0, 1)."""
    meas = np.atleast_2d(meas)
    pr1 = np.empty((modelparams.shape[0], meas.shape[0]))
    pr1[:, :] = np.einsum('ei,mi->me', meas, modelparams)
    np.clip(pr1, 0, 1, out=pr1)
    return pr1


# ---------------------------------------------------------------------------
# Model plugin objects (the surface SURVEY §8(b) says the updater consumes)
# ---------------------------------------------------------------------------

class _ModelBase(object):
    """abstract_model.py:96-117,274-281,357-395,443-468 — the members the hot path uses."""

    def __init__(self):
        self._call_count = 0

    @property
    def call_count(self):
        return self._call_count

    def _count(self, outcomes, modelparams, expparams):
        # abstract_model.py:466-468 (safe_shape = shape[0] or 1 for scalars)
        def ss(a):
            a = np.asarray(a)
            return a.shape[0] if a.ndim else 1
        self._call_count += ss(outcomes) * ss(modelparams) * ss(expparams)

    def are_models_valid(self, modelparams):
        return np.ones((modelparams.shape[0],), dtype=bool)

    def canonicalize(self, modelparams):
        return modelparams

    def update_timestep(self, modelparams, expparams):
        # abstract_model.py:374 — identity, as an (N, d, n_e) copy
        return np.tile(modelparams, (expparams.shape[0], 1, 1)).transpose((1, 2, 0))

    def clear_cache(self):
        pass

    def n_outcomes(self, expparams):
        return 2

    is_n_outcomes_constant = True

    @property
    def Q(self):
        # abstract_model.py:170-186: identity scale matrix unless a subclass says otherwise
        return np.ones((self.n_modelparams,))

    def domain(self, expparams):
        # abstract_model.py:287-298
        if expparams is None:
            return IntegerDomain(0, 1)
        n_o = np.broadcast_to(np.asarray(self.n_outcomes(expparams)), (np.asarray(expparams).shape[0],))
        return [IntegerDomain(0, int(k) - 1) for k in n_o]

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        """abstract_model.py:632-658 for constant two-outcome domains."""
        all_outcomes = np.arange(2)
        probabilities = self.likelihood(all_outcomes, modelparams, expparams)
        cdf = np.cumsum(probabilities, axis=0)
        randnum = np.random.random((repeat, 1, modelparams.shape[0], expparams.shape[0]))
        outcome_idxs = all_outcomes[np.argmax(cdf > randnum, axis=1)]
        outcomes = all_outcomes[outcome_idxs]
        if repeat == 1 and expparams.shape[0] == 1 and modelparams.shape[0] == 1:
            return outcomes[0, 0, 0]
        return outcomes


class IntegerDomain(object):
    """domains.py:427-560: the ``values`` of an integer outcome domain."""

    def __init__(self, min=0, max=1):
        self.min, self.max = int(min), int(max)

    @property
    def values(self):
        return np.arange(self.min, self.max + 1, dtype=int)


class CoinModel(_ModelBase):
    """test_models.py:262-326: pr0 is the model parameter itself."""
    n_modelparams = 1
    expparams_dtype = []

    def are_models_valid(self, modelparams):
        return np.logical_and(modelparams >= 0, modelparams <= 1).all(axis=1)     # test_models.py:303-304

    def likelihood(self, outcomes, modelparams, expparams):
        self._count(outcomes, modelparams, expparams)
        pr0 = np.tile(modelparams.flatten(), (expparams.shape[0], 1)).T           # test_models.py:323
        return two_outcome_likelihood(outcomes, pr0)


class SimpleInversionModel(_ModelBase):
    """test_models.py:64-166."""

    def __init__(self, min_freq=0):
        super(SimpleInversionModel, self).__init__()
        self._min_freq = min_freq

    n_modelparams = 1
    expparams_dtype = [('t', 'float'), ('w_', 'float')]

    def are_models_valid(self, modelparams):
        return np.all(modelparams > self._min_freq, axis=1)

    def likelihood(self, outcomes, modelparams, expparams):
        self._count(outcomes, modelparams, expparams)
        return two_outcome_likelihood(outcomes, precession_pr0(modelparams, expparams['t'], expparams['w_']))


class SimplePrecessionModel(SimpleInversionModel):
    """test_models.py:169-197 — scalar expparam t, w_ = 0."""
    expparams_dtype = 'float'

    def likelihood(self, outcomes, modelparams, expparams):
        self._count(outcomes, modelparams, expparams)
        expparams = np.asarray(expparams)
        t = expparams['t'] if expparams.dtype.names else expparams
        return two_outcome_likelihood(outcomes, precession_pr0(modelparams, t, 0))


class RandomizedBenchmarkingModel(_ModelBase):
    """rb.py:86-195 (zeroth order)."""

    def __init__(self, interleaved=False):
        super(RandomizedBenchmarkingModel, self).__init__()
        self._il = interleaved

    @property
    def n_modelparams(self):
        return 4 if self._il else 3

    @property
    def expparams_dtype(self):
        return [('m', 'uint')] + ([('reference', bool)] if self._il else [])

    def are_models_valid(self, modelparams):
        return rb_valid(modelparams, self._il)

    def likelihood(self, outcomes, modelparams, expparams):
        self._count(outcomes, modelparams, expparams)
        ref = expparams['reference'] if self._il else None
        return two_outcome_likelihood(outcomes, rb_pr0(modelparams, expparams['m'], self._il, ref))


class MLEModel(_ModelBase):
    """derived_models.py:681-703 — every likelihood of the underlying model raised to ``likelihood_power``."""

    def __init__(self, underlying_model, likelihood_power):
        super(MLEModel, self).__init__()
        self._underlying_model = underlying_model
        self._pow = likelihood_power

    underlying_model = property(lambda self: self._underlying_model)
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)
    is_n_outcomes_constant = property(lambda self: self._underlying_model.is_n_outcomes_constant)

    def n_outcomes(self, expparams):
        return self._underlying_model.n_outcomes(expparams)

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def canonicalize(self, modelparams):
        return self._underlying_model.canonicalize(modelparams)

    def update_timestep(self, modelparams, expparams):
        return self._underlying_model.update_timestep(modelparams, expparams)

    def likelihood(self, outcomes, modelparams, expparams):
        L = self._underlying_model.likelihood(outcomes, modelparams, expparams)      # derived_models.py:701-703
        return L ** self._pow


def binom_est_error(p, N, hedge=float(0)):
    # utils.py:683-688
    return np.sqrt(p * (1 - p) / (N + 2 * hedge + 1))


class PoisonedModel(_ModelBase):
    """derived_models.py:148-220 — the underlying likelihood plus clipped Gaussian noise that mimics the sampling
    error of adaptive (ALE: fixed tolerance) or fixed-sample (MLE: hedged binomial standard error) likelihood
    estimation."""

    def __init__(self, underlying_model, tol=None, n_samples=None, hedge=None):
        super(PoisonedModel, self).__init__()
        self._underlying_model = underlying_model
        if (tol is None) == (n_samples is None):
            raise ValueError("Exactly one of tol and n_samples must be specified")
        self._ale = tol is not None
        self._tol = tol
        self._n_samples = n_samples
        self._hedge = hedge if hedge is not None else 0.0

    underlying_model = property(lambda self: self._underlying_model)
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def likelihood(self, outcomes, modelparams, expparams):
        # derived_models.py:188-204
        L = self._underlying_model.likelihood(outcomes, modelparams, expparams)
        epsilon = np.random.normal(size=L.shape)
        if self._ale:
            epsilon *= self._tol
        else:
            epsilon *= binom_est_error(p=L, N=self._n_samples, hedge=self._hedge)
        np.clip(L + epsilon, 0, 1, out=L)
        return L


class RandomWalkModel(_ModelBase):
    """derived_models.py:705-741 — after every update each particle takes a step drawn from ``step_distribution``."""

    def __init__(self, underlying_model, step_distribution):
        super(RandomWalkModel, self).__init__()
        self._underlying_model = underlying_model
        self._step_dist = step_distribution
        if underlying_model.n_modelparams != step_distribution.n_rvs:
            raise TypeError("Step distribution does not match model dimension.")

    underlying_model = property(lambda self: self._underlying_model)
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def likelihood(self, outcomes, modelparams, expparams):
        return self._underlying_model.likelihood(outcomes, modelparams, expparams)

    def update_timestep(self, modelparams, expparams):
        # derived_models.py:733-741: one step per (particle, experiment), independent of the experiment
        steps = self._step_dist.sample(n=modelparams.shape[0] * expparams.shape[0])
        steps = steps.reshape((modelparams.shape[0], expparams.shape[0], self.n_modelparams))
        steps = steps.transpose((0, 2, 1))
        return modelparams[:, :, np.newaxis] + steps


class NormalStepDistribution(object):
    """distributions.py MultivariateNormalDistribution restricted to what RandomWalkModel needs: zero-mean steps
    ``np.random.multivariate_normal(mean, cov, n)`` (distributions.py:1016-1017)."""

    def __init__(self, mean, cov):
        self._mean, self._cov = np.asarray(mean, dtype=float), np.asarray(cov, dtype=float)
        self.n_rvs = self._mean.shape[0]

    def sample(self, n=1):
        return np.random.multivariate_normal(self._mean, self._cov, n)


class GaussianRandomWalkModel(_ModelBase):
    """derived_models.py:743-963, diagonal covariance without a model transformation: every update adds zero-mean
    Gaussian steps to the parameters ``random_walk_idxs``; the step scales are fixed (``fixed_covariance``, the
    diagonal) or learned (one extra model parameter sigma per walking parameter, valid iff >= 0)."""

    def __init__(self, underlying_model, random_walk_idxs='all', fixed_covariance=None, scale_mult=None):
        super(GaussianRandomWalkModel, self).__init__()
        self._underlying_model = underlying_model
        n_u = underlying_model.n_modelparams
        self._rw_idxs = np.s_[:n_u] if isinstance(random_walk_idxs, str) and random_walk_idxs == 'all' \
            else random_walk_idxs
        explicit = np.arange(n_u)[self._rw_idxs]
        if explicit.size == 0:
            raise IndexError('At least one model parameter must take a random walk.')
        self._n_rw = len(explicit)
        if fixed_covariance is None:                                   # derived_models.py:811-818
            self._has_fixed_covariance = False
            self._srw_idxs = (n_u + np.arange(self._n_rw)).astype(int)
            self._n_mps = n_u + self._n_rw
        else:                                                          # derived_models.py:834-842
            self._has_fixed_covariance = True
            fixed_covariance = np.asarray(fixed_covariance, dtype=float)
            if fixed_covariance.ndim != 1 or fixed_covariance.size != self._n_rw:
                raise ValueError('fixed_covariance must be the diagonal, one entry per walking parameter')
            self._fixed_scale = np.sqrt(fixed_covariance)
            self._n_mps = n_u
        if scale_mult is None:                                         # derived_models.py:859-864
            self._scale_mult_fcn = (lambda expparams: 1)
        elif isinstance(scale_mult, str):
            self._scale_mult_fcn = lambda x: x[scale_mult]
        else:
            self._scale_mult_fcn = scale_mult

    underlying_model = property(lambda self: self._underlying_model)
    n_modelparams = property(lambda self: self._n_mps)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)
    is_n_outcomes_constant = False

    def are_models_valid(self, modelparams):
        # derived_models.py:883-892
        n_u = self._underlying_model.n_modelparams
        ud_valid = self._underlying_model.are_models_valid(modelparams[..., :n_u])
        if self._has_fixed_covariance:
            return ud_valid
        pos_std = np.greater_equal(modelparams[..., self._srw_idxs], 0).all(axis=-1)
        return np.logical_and(ud_valid, pos_std)

    def likelihood(self, outcomes, modelparams, expparams):
        # derived_models.py:894-896
        return self._underlying_model.likelihood(
            outcomes, modelparams[..., :self._underlying_model.n_modelparams], expparams)

    def update_timestep(self, modelparams, expparams):
        # derived_models.py:921-963 (diagonal branch, no transformation)
        n_mps, n_eps = modelparams.shape[0], expparams.shape[0]
        scale = self._fixed_scale if self._has_fixed_covariance else modelparams[:, self._srw_idxs]
        steps = scale * np.random.normal(size=(n_eps, n_mps, self._n_rw))
        steps = steps.transpose((1, 2, 0))
        steps = self._scale_mult_fcn(expparams) * steps
        new_mps = np.repeat(modelparams[:, :, np.newaxis], n_eps, axis=2)
        new_mps[:, self._rw_idxs, :] += steps
        return new_mps


class BinomialModel(_ModelBase):
    """derived_models.py:222-360 — n_meas iid shots of a two-outcome model."""

    is_n_outcomes_constant = False

    def __init__(self, underlying_model):
        super(BinomialModel, self).__init__()
        self._underlying_model = underlying_model
        if isinstance(underlying_model.expparams_dtype, str):
            self._expparams_scalar = True
            self._expparams_dtype = [('x', underlying_model.expparams_dtype), ('n_meas', 'uint')]
        else:
            self._expparams_scalar = False
            self._expparams_dtype = underlying_model.expparams_dtype + [('n_meas', 'uint')]

    underlying_model = property(lambda self: self._underlying_model)
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    expparams_dtype = property(lambda self: self._expparams_dtype)

    def n_outcomes(self, expparams):
        return expparams['n_meas'] + 1

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def canonicalize(self, modelparams):
        return self._underlying_model.canonicalize(modelparams)

    def _pr1(self, modelparams, expparams):
        return self._underlying_model.likelihood(
            np.array([1], dtype='uint'), modelparams,
            expparams['x'] if self._expparams_scalar else expparams)

    def likelihood(self, outcomes, modelparams, expparams):
        # derived_models.py:314-329
        self._count(outcomes, modelparams, expparams)
        outcomes = np.atleast_1d(np.asarray(outcomes))
        pr1 = self._pr1(modelparams, expparams)
        L = np.concatenate([
            binomial_pmf(expparams['n_meas'][np.newaxis, :], outcomes[i], pr1)
            for i in range(outcomes.shape[0])])
        assert not np.any(np.isnan(L))
        return L

    def update_timestep(self, modelparams, expparams):
        return self._underlying_model.update_timestep(
            modelparams, expparams['x'] if self._expparams_scalar else expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        # derived_models.py:331-355
        pr1 = self._pr1(modelparams, expparams)
        dist = scipy.stats.binom(expparams['n_meas'].astype('int'), pr1[0, :, :])
        if pr1.size != 1:
            os_ = np.concatenate([dist.rvs()[np.newaxis, :, :] for _ in range(repeat)], axis=0)
        else:
            os_ = np.concatenate([np.array([[[dist.rvs()]]]) for _ in range(repeat)], axis=0)
        return os_[0, 0, 0] if os_.size == 1 else os_


def gell_mann_basis_data(dim):
    """tomography/bases.py:71-111 — generalised Gell-Mann matrices, (dim^2, dim, dim)."""
    B = np.zeros((dim ** 2, dim, dim), dtype=complex)
    B[0] = np.eye(dim) / np.sqrt(dim)
    for k in range(1, dim):
        diag = np.concatenate([np.ones((k,)), [-k], np.zeros((dim - k - 1,))])
        B[k] = np.diag(diag) / np.sqrt(k + k ** 2)
    off = dim * (dim - 1) // 2
    for i in range(1, dim):
        for j in range(i):
            k = (i - 1) * i // 2 + j + dim
            B[k, [i, j], [j, i]] = 1 / np.sqrt(2)
            B[k + off, [i, j], [j, i]] = [1j / np.sqrt(2), -1j / np.sqrt(2)]
    return B


def pauli_basis_data(nq=1):
    """tomography/bases.py:113-153 — nq-fold tensor product of {1, X, Y, Z}/sqrt(2)."""
    single = gell_mann_basis_data(2)[[0, 2, 3, 1]]
    dim = 2 ** nq
    out = np.zeros((dim ** 2, dim, dim), dtype=complex)
    for idx, factors in enumerate(itertools.product(*([single] * nq))):
        out[idx] = reduce(np.kron, factors)
    return out


class TomographyBasis(object):
    """tomography/bases.py:180-321 — only ``data``, ``dim`` and ``flat()``."""

    def __init__(self, data):
        self.data = data
        self.dim = data.shape[1]
        self._flat = data.reshape((data.shape[0], -1))

    def flat(self):
        return self._flat


def pauli_basis(nq=1):
    return TomographyBasis(pauli_basis_data(nq))


def gell_mann_basis(dim):
    return TomographyBasis(gell_mann_basis_data(dim))


class TomographyModel(_ModelBase):
    """tomography/models.py:82-226."""

    def __init__(self, basis, allow_subnormalized=False):
        super(TomographyModel, self).__init__()
        self._dim = basis.dim
        self._basis = basis
        self._allow_subnormalied = allow_subnormalized

    @property
    def n_modelparams(self):
        return self._dim ** 2

    @property
    def expparams_dtype(self):
        return [('meas', float, self._dim ** 2)]

    def trunc_neg_eigs(self, particle):
        # tomography/models.py:172-192
        arr = np.tensordot(particle, self._basis.data.conj(), 1)
        w, v = np.linalg.eig(arr)
        if np.all(w >= 0):
            return particle
        w[w < 0] = 0
        new_arr = np.dot(v * w, v.conj().T)
        new_particle = np.real(np.dot(self._basis.flat(), new_arr.flatten()))
        assert new_particle[0] > 0
        return new_particle

    def renormalize(self, modelparams):
        # tomography/models.py:194-209
        norm = modelparams[:, 0] * np.sqrt(self._dim)
        assert not np.sum(norm == 0)
        return modelparams / norm[:, None]

    def canonicalize(self, modelparams):
        # tomography/models.py:149-170
        modelparams = np.apply_along_axis(self.trunc_neg_eigs, 1, modelparams)
        if not self._allow_subnormalied:
            modelparams = self.renormalize(modelparams)
        return modelparams

    def likelihood(self, outcomes, modelparams, expparams):
        self._count(outcomes, modelparams, expparams)
        return two_outcome_likelihood(outcomes, 1 - tomography_pr1(modelparams, expparams['meas']))


class DiffusiveTomographyModel(TomographyModel):
    """tomography/models.py:228-272 — one extra model parameter eps > 0 (diffusion scale) and one extra experiment
    field ``t``; every update ends with a Gaussian step of scale eps * sqrt(t) on the state parameters 1 .. d^2 - 1
    followed by re-canonicalisation."""

    @property
    def n_modelparams(self):
        return self._dim ** 2 + 1

    @property
    def expparams_dtype(self):
        return [('meas', float, self._dim ** 2), ('t', float)]

    def are_models_valid(self, modelparams):
        # tomography/models.py:245-249
        return np.logical_and(np.ones((modelparams.shape[0],), dtype=bool), modelparams[:, -1] > 0)

    def canonicalize(self, modelparams):
        # tomography/models.py:251-255
        return np.concatenate([
            super(DiffusiveTomographyModel, self).canonicalize(modelparams[:, :-1]),
            modelparams[:, -1, None]
        ], axis=1)

    def likelihood(self, outcomes, modelparams, expparams):
        # tomography/models.py:256-257
        return super(DiffusiveTomographyModel, self).likelihood(outcomes, modelparams[:, :-1], expparams)

    def update_timestep(self, modelparams, expparams):
        # tomography/models.py:259-272
        eps = (modelparams[:, -1, None] * np.sqrt(expparams['t']))[:, :, None]
        steps = eps * np.random.randn(*modelparams[:, None, :].shape)
        steps[:, :, [0, -1]] = 0
        raw_modelparams = modelparams[:, None, :] + steps
        for idx_experiment in range(len(expparams)):
            raw_modelparams[:, idx_experiment, :] = self.canonicalize(raw_modelparams[:, idx_experiment, :])
        return raw_modelparams.transpose((0, 2, 1))


# ---------------------------------------------------------------------------
# Priors used by the benchmark configurations (host side)
# ---------------------------------------------------------------------------

class UniformDistribution(object):
    """distributions.py:792-827."""

    def __init__(self, ranges=np.array([[0, 1]])):
        ranges = np.asarray(ranges, dtype=float)
        if ranges.ndim == 1:
            ranges = ranges[np.newaxis, ...]
        self._ranges = ranges
        self._n_rvs = ranges.shape[0]
        self._delta = ranges[:, 1] - ranges[:, 0]

    n_rvs = property(lambda self: self._n_rvs)

    def sample(self, n=1):
        shape = (n, self._n_rvs)
        z = np.random.random(shape)
        return self._ranges[:, 0] + self._delta * z


class PostselectedDistribution(object):
    """distributions.py:1304-1350 — redraw until the model says valid."""

    def __init__(self, distribution, model, maxiters=100):
        self._dist = distribution
        self._model = model
        self._maxiters = maxiters

    n_rvs = property(lambda self: self._dist.n_rvs)

    def sample(self, n=1):
        samples = np.empty((n, self.n_rvs))
        idxs_to_sample = np.arange(n)
        iters = 0
        while idxs_to_sample.size and iters < self._maxiters:
            samples[idxs_to_sample] = self._dist.sample(len(idxs_to_sample))
            idxs_to_sample = idxs_to_sample[np.nonzero(np.logical_not(
                self._model.are_models_valid(samples[idxs_to_sample, :])))[0]]
            iters += 1
        if idxs_to_sample.size:
            raise RuntimeError("Did not successfully postselect within {} iterations.".format(self._maxiters))
        return samples


class GinibreTomographyPrior(object):
    """Restated Ginibre prior (the reference's needs QuTiP, absent here):
    X = randn + i randn (dim x dim), rho = X X^H / tr, x_a = Re tr(B_a^* rho)
    (tomography/distributions.py:138-141,193-196; bases.py:323-336)."""

    def __init__(self, basis):
        self._basis = basis
        self._dim = basis.dim

    n_rvs = property(lambda self: self._dim ** 2)

    def sample(self, n=1):
        d = self._dim
        X = np.random.randn(n, d, d) + 1j * np.random.randn(n, d, d)
        rho = np.einsum('nij,nkj->nik', X, X.conj())
        rho /= np.trace(rho, axis1=1, axis2=2)[:, None, None]
        return np.real(np.einsum('aij,nij->na', self._basis.data.conj(), rho))


# ---------------------------------------------------------------------------
# Particle container
# ---------------------------------------------------------------------------

def n_ess(weights):
    """distributions.py:299-307."""
    return 1 / (np.sum(weights ** 2))


def particle_mean(weights, locations):
    """distributions.py:337-348."""
    return np.dot(weights, locations)


def particle_covariance_mtx(weights, locations):
    """distributions.py:351-399 — E[x x^T] - mu mu^T, then the PSD warning."""
    mu = particle_mean(weights, locations)
    xs = locations.transpose([1, 0])
    cov = np.einsum('i,mi,ni', weights, xs, xs) - np.dot(mu[..., np.newaxis], mu[np.newaxis, ...])
    assert np.all(np.isfinite(cov))
    if not np.all(scipy.linalg.eig(cov)[0] >= 0):
        warnings.warn('Numerical error in covariance estimation causing positive semidefinite violation.',
                      ApproximationWarning)
    return cov


def sqrtm_psd(A, est_error=True, check_finite=True):
    """utils.py:593-607."""
    w, v = scipy.linalg.eigh(A, check_finite=check_finite)
    mask = w <= 0
    w[mask] = 0
    np.sqrt(w, out=w)
    A_sqrt = (v * w).dot(v.conj().T)
    if est_error:
        return A_sqrt, np.linalg.norm(np.dot(A_sqrt, A_sqrt) - A, 'fro')
    return A_sqrt


class ParticleDistribution(object):
    """distributions.py:261-453 (constructor, n_ess, moments, sample)."""

    def __init__(self, n_mps=None, particle_locations=None, particle_weights=None):
        if particle_locations is None or particle_weights is None:
            self.particle_locations = np.zeros((1, n_mps))
            self.particle_weights = np.ones((1,))
        elif n_mps is None:
            self.particle_locations = particle_locations
            self.particle_weights = np.abs(particle_weights)
            self.particle_weights = self.particle_weights / np.sum(self.particle_weights)
        else:
            raise ValueError('Either the dimension of parameter space, `n_mps`, or the particles, '
                             '`particle_locations` and `particle_weights` must be specified.')

    n_particles = property(lambda self: self.particle_locations.shape[0])
    n_rvs = property(lambda self: self.particle_locations.shape[1])
    n_ess = property(lambda self: n_ess(self.particle_weights))

    def sample(self, n=1):
        cdf = np.cumsum(self.particle_weights)
        return self.particle_locations[np.minimum(
            cdf.searchsorted(np.random.random((n,)), side='right'), len(cdf) - 1)]

    def est_mean(self):
        return particle_mean(self.particle_weights, self.particle_locations)

    def est_meanfn(self, fn):
        return np.einsum('i...,i...', self.particle_weights, fn(self.particle_locations))

    def est_covariance_mtx(self, corr=False):
        cov = particle_covariance_mtx(self.particle_weights, self.particle_locations)
        if corr:
            dstd = np.sqrt(np.diag(cov))
            cov /= np.outer(dstd, dstd)
        return cov


# ---------------------------------------------------------------------------
# Liu-West resampler
# ---------------------------------------------------------------------------

def liu_west_draw(weights, n_particles):
    """resamplers.py:308-321 — sequential CDF, legacy uniforms, right-bisect.
    Returns (js, cdf, u).  NOT clamped, exactly like the reference."""
    cdf = np.cumsum(weights)
    u = np.random.random((n_particles,))
    return cdf.searchsorted(u, side='right'), cdf, u


class LiuWestResampler(object):
    """resamplers.py:171-392."""

    _override_h = False

    def __init__(self, a=0.98, h=None, maxiter=1000, debug=False, postselect=True,
                 zero_cov_comp=1e-10, default_n_particles=None, kernel=np.random.randn):
        self._default_n_particles = default_n_particles
        self.a = a
        if h is not None:
            self._override_h = True
            self._h = h
        self._maxiter = maxiter
        self._debug = debug
        self._postselect = postselect
        self._zero_cov_comp = zero_cov_comp
        self._kernel = kernel
        self.trace = None   # filled with intermediates when ``record`` is set
        self.record = False

    @property
    def a(self):
        return self._a

    @a.setter
    def a(self, new_a):
        self._a = new_a
        if not self._override_h:
            self._h = np.sqrt(1 - new_a ** 2)

    @property
    def h(self):
        return self._h

    def __call__(self, model, particle_dist, n_particles=None,
                 precomputed_mean=None, precomputed_cov=None):
        mean = particle_dist.est_mean() if precomputed_mean is None else precomputed_mean
        cov = particle_dist.est_covariance_mtx() if precomputed_cov is None else precomputed_cov
        if n_particles is None:
            n_particles = (particle_dist.n_particles if self._default_n_particles is None
                           else self._default_n_particles)
        a, h = self._a, self._h
        if scipy.linalg.norm(cov, 'fro') == 0:
            warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                          "Consider increasing n_particles to improve covariance estimates.",
                          ResamplerWarning)
            cov = self._zero_cov_comp * np.eye(cov.shape[0])
        S, S_err = sqrtm_psd(cov)
        if not np.isfinite(S_err):
            raise ResamplerError("Infinite error in computing the square root of the "
                                 "covariance matrix. Check that n_ess is not too small.")
        S = np.real(h * S)

        w = particle_dist.particle_weights
        l = particle_dist.particle_locations
        n_rvs = particle_dist.n_rvs

        new_locs = np.empty((n_particles, n_rvs))
        js, cdf, u = liu_west_draw(w, n_particles)
        js0 = js
        idxs = np.arange(n_particles, dtype=int)
        mus = a * l[js, :] + (1 - a) * mean       # resamplers.py:325

        normals = []
        invalid_history = []
        n_iters = 0
        while idxs.size and n_iters < self._maxiter:
            n_iters += 1
            eps = self._kernel(n_rvs, mus.shape[0])
            if self.record:
                normals.append(eps.copy())
            new_locs[idxs, :] = mus + np.dot(S, eps).T          # :332
            cand = new_locs[idxs, :]
            if self._postselect:
                valid = model.are_models_valid(cand)
            else:
                valid = np.ones((cand.shape[0],), dtype=bool)
            idxs = idxs[np.nonzero(np.logical_not(valid))[0]]   # :365-367
            js = js[np.logical_not(valid)]                      # :371
            mus = mus[:idxs.size, :]                            # :372 (prefix slice quirk)
            if self.record:
                invalid_history.append(idxs.copy())
        if idxs.size:
            warnings.warn(("Liu-West resampling failed to find valid models for {} "
                           "particles within {} iterations.").format(idxs.size, self._maxiter),
                          ResamplerWarning)
        if self.record:
            self.trace = dict(mean=mean, cov=cov, S=S, cdf=cdf, u=u, js=js0, normals=normals,
                              invalid=invalid_history, n_iters=n_iters)
        return ParticleDistribution(particle_locations=new_locs,
                                    particle_weights=np.ones((n_particles,)) / n_particles)


# ---------------------------------------------------------------------------
# SMC updater
# ---------------------------------------------------------------------------

class SMCUpdater(ParticleDistribution):
    """smc.py:97-551 — the hot-path members only."""

    def __init__(self, model, n_particles, prior, resample_a=None, resampler=None,
                 resample_thresh=0.5, debug_resampling=False, track_resampling_divergence=False,
                 zero_weight_policy='error', zero_weight_thresh=None, canonicalize=True):
        super(SMCUpdater, self).__init__(particle_locations=np.zeros((0, model.n_modelparams)),
                                         particle_weights=np.zeros((0,)))
        self._resample_count = 0
        self._min_n_ess = n_particles
        self.model = model
        self.prior = prior
        self._canonicalize = bool(canonicalize)
        if resample_a is not None:
            warnings.warn("The 'resample_a' keyword argument is deprecated; use "
                          "'resampler=LiuWestResampler(a)' instead.", DeprecationWarning)
            if resampler is not None:
                raise ValueError("Both a resample_a and an explicit resampler were provided; "
                                 "please provide only one.")
            self.resampler = LiuWestResampler(a=resample_a)
        elif resampler is None:
            self.resampler = LiuWestResampler(default_n_particles=n_particles)
        else:
            self.resampler = resampler
        self.resample_thresh = resample_thresh
        self._just_resampled = False
        self._data_record = []
        self._normalization_record = []
        self._zero_weight_policy = zero_weight_policy
        self._zero_weight_thresh = (zero_weight_thresh if zero_weight_thresh is not None
                                    else 10 * np.spacing(1))
        self.reset(n_particles)

    resample_count = property(lambda self: self._resample_count)
    just_resampled = property(lambda self: self._just_resampled)
    normalization_record = property(lambda self: self._normalization_record)
    min_n_ess = property(lambda self: self._min_n_ess)
    data_record = property(lambda self: self._data_record[:])

    @property
    def log_total_likelihood(self):
        return np.sum(np.log(self.normalization_record))

    def reset(self, n_particles=None, only_params=None, reset_weights=True):
        # smc.py:281-320
        if n_particles is not None and only_params is not None:
            raise ValueError("Cannot set both n_particles and only_params.")
        if n_particles is None:
            n_particles = self.n_particles
        if reset_weights:
            self.particle_weights = np.ones((n_particles,)) / n_particles
        if only_params is None:
            sl = np.s_[:, :]
            self.particle_locations = np.zeros((n_particles, self.model.n_modelparams))
        else:
            sl = np.s_[:, only_params]
        self.particle_locations[sl] = self.prior.sample(n=n_particles)[sl]
        if self._canonicalize:
            self.particle_locations[sl] = self.model.canonicalize(self.particle_locations[sl])

    def hypothetical_update(self, outcomes, expparams, return_likelihood=False,
                            return_normalization=False):
        # smc.py:324-386
        weights = self.particle_weights
        locs = self.particle_locations
        if not isinstance(outcomes, np.ndarray):
            outcomes = np.array([outcomes])
        L = self.model.likelihood(outcomes, locs, expparams).transpose([0, 2, 1])
        hyp_weights = weights * L
        norm_scale = np.sum(hyp_weights, axis=2)[..., np.newaxis]
        fixed_norm_scale = norm_scale.copy()
        fixed_norm_scale[np.abs(norm_scale) < np.spacing(1)] = 1
        norm_weights = hyp_weights / fixed_norm_scale
        out = (norm_weights,)
        if return_likelihood:
            out += (L,)
        if return_normalization:
            out += (norm_scale,)
        return out[0] if len(out) == 1 else out

    def _hypothetical_posteriors(self, expparams):
        # smc.py:576-595 / 628-647: every outcome's hypothetical weights and normalisation; the last outcome's
        # likelihood is the complement of the others
        os_ = self.model.domain(expparams[0, np.newaxis])[0].values
        w_hyp, L, N = self.hypothetical_update(os_[:-1], expparams, return_normalization=True,
                                               return_likelihood=True)
        w_hyp_last_outcome = (1 - L.sum(axis=0)) * self.particle_weights[np.newaxis, :]
        N = np.concatenate([N[:, :, 0], np.sum(w_hyp_last_outcome[np.newaxis, :, :], axis=2)], axis=0)
        w_hyp_last_outcome = w_hyp_last_outcome / N[-1, :, np.newaxis]
        w_hyp = np.concatenate([w_hyp, w_hyp_last_outcome[np.newaxis, :, :]], axis=0)
        return w_hyp, N

    def bayes_risk(self, expparams):
        # smc.py:553-605
        n_eps = expparams.size
        if n_eps > 1 and not self.model.is_n_outcomes_constant:
            risk = np.empty(n_eps)
            for idx in range(n_eps):
                risk[idx] = self.bayes_risk(expparams[idx, np.newaxis])[0]
            return risk
        w_hyp, N = self._hypothetical_posteriors(expparams)
        mu_hyp = np.dot(w_hyp, self.particle_locations)
        var_hyp = np.sum(
            w_hyp * np.sum(self.model.Q * (self.particle_locations[np.newaxis, np.newaxis, :, :]
                                           - mu_hyp[:, :, np.newaxis, :]) ** 2, axis=3),
            axis=2)
        return np.sum(N * var_hyp, axis=0)

    def expected_information_gain(self, expparams):
        # smc.py:607-657
        n_eps = expparams.size
        if n_eps > 1 and not self.model.is_n_outcomes_constant:
            gain = np.empty(n_eps)
            for idx in range(n_eps):
                gain[idx] = self.expected_information_gain(expparams[idx, np.newaxis])[0]
            return gain
        w_hyp, N = self._hypothetical_posteriors(expparams)
        KLD = np.sum(w_hyp * np.log(w_hyp / self.particle_weights), axis=2)
        return np.sum(N * KLD, axis=0)

    def update(self, outcome, expparams, check_for_resample=True):
        # smc.py:388-457
        self._data_record.append(outcome)
        self._just_resampled = False
        weights, norm = self.hypothetical_update(outcome, expparams, return_normalization=True)
        if not np.all(weights >= 0):
            warnings.warn("Negative weights occured in particle approximation. Smallest weight "
                          "observed == {}. Clipping weights.".format(np.min(weights)),
                          ApproximationWarning)
            np.clip(weights, 0, 1, out=weights)
        if np.sum(weights) <= self._zero_weight_thresh:
            policy = self._zero_weight_policy
            if policy == 'ignore':
                pass
            elif policy == 'skip':
                return
            elif policy == 'warn':
                warnings.warn("All particle weights are zero. This will very likely fail quite badly.",
                              ApproximationWarning)
            elif policy == 'error':
                raise RuntimeError("All particle weights are zero.")
            elif policy == 'reset':
                warnings.warn("All particle weights are zero. Resetting from initial prior.",
                              ApproximationWarning)
                self.reset()
            else:
                raise ValueError("Invalid zero-weight policy {} encountered.".format(policy))
        self.particle_weights[:] = weights[0, 0, :]
        self._normalization_record.append(norm[0][0])
        self.particle_locations = self.model.update_timestep(self.particle_locations, expparams)[:, :, 0]
        if self.n_ess <= self._min_n_ess:
            self._min_n_ess = self.n_ess
        if check_for_resample:
            self._maybe_resample()

    def batch_update(self, outcomes, expparams, resample_interval=5):
        # smc.py:459-487
        n_exps = outcomes.shape[0]
        if expparams.shape[0] != n_exps:
            raise ValueError("The number of outcomes and experiments must match.")
        if len(expparams.shape) == 1:
            expparams = expparams[:, None]
        for idx_exp, (outcome, experiment) in enumerate(zip(iter(outcomes), iter(expparams))):
            self.update(outcome, experiment, check_for_resample=False)
            if (idx_exp + 1) % resample_interval == 0:
                self._maybe_resample()

    def _maybe_resample(self):
        # smc.py:263-277
        ess = self.n_ess
        if ess <= 10:
            warnings.warn("Extremely small n_ess encountered ({}). Resampling is likely to fail. "
                          "Consider adding particles, or resampling more often.".format(ess),
                          ApproximationWarning)
        if ess < self.n_particles * self.resample_thresh:
            self.resample()

    def resample(self):
        # smc.py:491-551 (without the debug / divergence-tracking branches)
        if self.just_resampled:
            warnings.warn("Resampling without additional data; this may not perform as desired.",
                          ResamplerWarning)
        self._just_resampled = True
        self._resample_count += 1
        new_distribution = self.resampler(self.model, self)
        self.particle_weights = new_distribution.particle_weights
        self.particle_locations = new_distribution.particle_locations
        if self._canonicalize:
            self.particle_locations[:, :] = self.model.canonicalize(self.particle_locations)
        try:
            self.model.clear_cache()
        except Exception as e:  # pragma: no cover
            warnings.warn("Exception raised when clearing model cache: {}. Ignoring.".format(e))
