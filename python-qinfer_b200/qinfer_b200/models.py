"""Model plugin surface (host side).

Two things live here:

* ``describe_model`` — recognises the built-in model families of the hot path
   on ANY object exposing QInfer's Model interface — the real
  ``qinfer.SimplePrecessionModel`` etc. or the classes below — and turns it into
  the ``qb_model`` descriptor the kernels take.  Anything else raises
  ``UnsupportedModelError``: there is no CPU fallback.
* Model classes with the reference's names, constructor arguments and methods
  (``likelihood``, ``are_models_valid``, ``canonicalize``, ``n_modelparams``,
  ``expparams_dtype`` …) whose arithmetic runs on the GPU through the C ABI, so
  that the engine is usable where QInfer itself is not installed.
"""
import itertools
from functools import reduce

import numpy as np

from . import _lib
from ._exceptions import UnsupportedModelError


class ModelDescriptor(object):
    """What the kernels need to know about a model instance."""

    def __init__(self, kind, d, binomial=False, interleaved=False, min_freq=0.0,
                 scalar_expparam=False, binomial_scalar=False, dim=0, basis=None, allow_subnormalized=False,
                 likelihood_power=1.0, d_extra=0, extra_rule=0, walk=None, poison=None):
        self.kind = kind
        self.d = int(d)
        self.d_extra = int(d_extra)                 # trailing decorator parameters (learned walk scales, diffusion rate)
        self.extra_rule = int(extra_rule)
        self.walk = walk                            # time-dependent decorator: how update_timestep moves the particles
        self.poison = poison                        # PoisonedModel: {'mode': 0 ALE | 1 MLE, 'tol', 'denom'}
        self.binomial = bool(binomial)
        self.interleaved = bool(interleaved)
        self.min_freq = float(min_freq)
        self.scalar_expparam = scalar_expparam      # SimplePrecessionModel: expparams is a bare float array
        self.binomial_scalar = binomial_scalar      # BinomialModel renamed the scalar field to 'x'
        self.dim = int(dim)
        self.basis = basis
        self.allow_subnormalized = bool(allow_subnormalized)
        self.likelihood_power = float(likelihood_power)
        self.c_model = _lib.QbModel(kind=kind, d=self.d, binomial=int(self.binomial),
                                    interleaved=int(self.interleaved), min_freq=self.min_freq,
                                    likelihood_power=self.likelihood_power, d_extra=self.d_extra,
                                    extra_rule=self.extra_rule)

    @property
    def needs_canonicalize(self):
        return self.kind == _lib.QB_MODEL_TOMOGRAPHY

    def expparams_record(self, expparams, idx=0):
        """One element of an ``expparams`` array -> a new ``qb_expparams``."""
        return self.fill_record(_lib.QbExpparams(), expparams, idx)

    def fill_record(self, ep, expparams, idx=0):
        """Fill an existing ``qb_expparams`` (the hot loop re-uses one struct per updater)."""
        if self.kind == _lib.QB_MODEL_PRECESSION and not self.binomial and type(expparams) is np.ndarray \
                and expparams.dtype.names is None:
            ep.t = float(expparams.flat[idx])                 # SimplePrecessionModel: bare float array
            ep.w_ = 0.0
            return ep
        if self.kind == _lib.QB_MODEL_COIN and not self.binomial:
            return ep                                         # CoinModel: expparams_dtype is empty
        arr = np.asarray(expparams)
        names = arr.dtype.names
        rec = arr.reshape(-1)[idx] if arr.ndim else arr[()]
        inner = rec
        if self.binomial:
            if not names or 'n_meas' not in names:
                raise ValueError("BinomialModel expparams need an 'n_meas' field (derived_models.py:243-249)")
            ep.n_meas = int(rec['n_meas'])
            if self.binomial_scalar:
                inner = rec['x']
                names_inner = None
            else:
                names_inner = names
        else:
            names_inner = names
        if self.kind == _lib.QB_MODEL_PRECESSION:
            if names_inner and 't' in names_inner:
                ep.t = float(inner['t'])
                ep.w_ = float(inner['w_']) if 'w_' in names_inner and not self.scalar_expparam else 0.0
            else:
                ep.t = float(inner)
                ep.w_ = 0.0
        elif self.kind == _lib.QB_MODEL_RB:
            ep.m = int(inner['m'])
            ep.reference = int(bool(inner['reference'])) if self.interleaved else 0
        elif self.kind == _lib.QB_MODEL_COIN:
            pass
        else:
            meas = np.asarray(inner['meas'], dtype=float).reshape(-1)
            d_state = self.d - self.d_extra
            if meas.shape[0] != d_state:
                raise ValueError("expparams['meas'] has %d entries, model has %d state parameters"
                                 % (meas.shape[0], d_state))
            for c in range(d_state):
                ep.meas[c] = meas[c]
            for c in range(d_state, self.d):
                ep.meas[c] = 0.0                               # decorator parameters do not enter <meas, x>
            if names_inner and 't' in names_inner:
                ep.t = float(inner['t'])                       # DiffusiveTomographyModel: the time the step diffuses for
        return ep


def _name(obj):
    return type(obj).__name__


def describe_model(model):
    """Recognise a built-in model (by exact class name, so subclasses with their
    own likelihood are NOT silently mis-evaluated) and build its descriptor."""
    binomial = False
    binomial_scalar = False
    power = 1.0
    walk = poison = None
    d_extra = extra_rule = 0
    # f4 decorators (derived_models.py:148-220, 705-963), outermost first, each at most once
    for _ in range(2):
        if _name(model) == 'PoisonedModel' and poison is None:
            if getattr(model, '_tol', None) is not None:                     # ALE: fixed tolerance
                poison = {'mode': 0, 'tol': float(model._tol), 'denom': 1.0}
            else:                                                            # MLE: hedged binomial standard error
                hedge = getattr(model, '_hedge', 0.0) or 0.0
                poison = {'mode': 1, 'tol': 0.0, 'denom': float(model._n_samples + 2 * hedge + 1)}
            model = model.underlying_model
        elif _name(model) == 'RandomWalkModel' and walk is None:
            walk = {'kind': 'generic', 'dist': model._step_dist}
            model = model.underlying_model
        elif _name(model) == 'GaussianRandomWalkModel' and walk is None:
            walk, d_extra, extra_rule = _describe_gaussian_walk(model)
            model = model.underlying_model
    inner = model
    if _name(model) == 'MLEModel':            # derived_models.py:681-703: L ** likelihood_power, outermost decorator
        power = float(getattr(model, '_pow'))
        model = inner = model.underlying_model
    if _name(model) == 'BinomialModel':
        binomial = True
        binomial_scalar = bool(getattr(model, '_expparams_scalar'))
        inner = model.underlying_model
    name = _name(inner)
    extra = dict(d_extra=d_extra, extra_rule=extra_rule, walk=walk, poison=poison)
    if walk is not None and walk['kind'] != 'generic':
        n_under = int(walk.pop('n_under'))
        if n_under != int(inner.n_modelparams):
            raise UnsupportedModelError("GaussianRandomWalkModel over a %d-parameter chain, expected %d"
                                        % (int(inner.n_modelparams), n_under))
    if name in ('SimplePrecessionModel', 'SimpleInversionModel'):
        return ModelDescriptor(_lib.QB_MODEL_PRECESSION, 1 + d_extra, binomial=binomial, **extra,
                               min_freq=getattr(inner, '_min_freq', 0.0),
                               scalar_expparam=(name == 'SimplePrecessionModel'),
                               binomial_scalar=binomial_scalar, likelihood_power=power)
    if name == 'RandomizedBenchmarkingModel':
        il = bool(getattr(inner, '_il', False))
        return ModelDescriptor(_lib.QB_MODEL_RB, (4 if il else 3) + d_extra, binomial=binomial, interleaved=il, **extra,
                               binomial_scalar=binomial_scalar, likelihood_power=power)
    if name == 'CoinModel':
        return ModelDescriptor(_lib.QB_MODEL_COIN, 1 + d_extra, binomial=binomial, binomial_scalar=binomial_scalar,
                               likelihood_power=power, **extra)
    if name in ('TomographyModel', 'DiffusiveTomographyModel'):
        if name == 'DiffusiveTomographyModel':
            if walk is not None or d_extra:
                raise UnsupportedModelError("a random-walk decorator over DiffusiveTomographyModel is not supported")
            # tomography/models.py:228-272: one extra parameter eps > 0, the per-unit-time diffusion scale
            extra.update(d_extra=1, extra_rule=2, walk={'kind': 'diffusive'})
        dim = int(getattr(inner, '_dim'))
        basis = np.ascontiguousarray(np.asarray(inner._basis.data, dtype=complex))
        if dim ** 2 > _lib.QB_MAX_D:
            raise UnsupportedModelError("TomographyModel with dim=%d exceeds QB_MAX_D" % dim)
        return ModelDescriptor(_lib.QB_MODEL_TOMOGRAPHY, dim ** 2 + extra['d_extra'], binomial=binomial, dim=dim,
                               basis=basis, allow_subnormalized=getattr(inner, '_allow_subnormalied', False),
                               binomial_scalar=binomial_scalar, likelihood_power=power, **extra)
    raise UnsupportedModelError(
        "%s is not one of the model families the B200 kernels implement (SimplePrecessionModel, "
        "SimpleInversionModel, RandomizedBenchmarkingModel, CoinModel, tomography.TomographyModel / "
        "DiffusiveTomographyModel, optionally wrapped in BinomialModel, MLEModel, PoisonedModel, RandomWalkModel or "
        "GaussianRandomWalkModel). There is no CPU fallback." % name)


def _describe_gaussian_walk(model):
    """GaussianRandomWalkModel (derived_models.py:743-963) -> (walk description, d_extra, extra_rule)."""
    if getattr(model, '_has_transformation', False):
        raise UnsupportedModelError("GaussianRandomWalkModel with a model_transformation is not implemented on the "
                                    "device (it is an arbitrary host callable)")
    under = model.underlying_model
    n_under = int(under.n_modelparams)
    idxs = np.arange(n_under)[model._rw_idxs].astype(int).reshape(-1)
    scale_mult = getattr(model, '_scale_mult_fcn', None)
    walk = {'idxs': [int(i) for i in idxs], 'scale_mult': scale_mult, 'n_under': n_under}
    diagonal = bool(getattr(model, '_diagonal', True))
    if getattr(model, '_has_fixed_covariance'):
        if diagonal:
            walk.update(kind='fixed', scale=[float(v) for v in np.asarray(model._fixed_scale).reshape(-1)])
        else:
            walk.update(kind='fixed_dense', chol=np.asarray(model._fixed_chol, dtype=float))
        return walk, 0, 0
    if not diagonal:
        raise UnsupportedModelError("GaussianRandomWalkModel with a learned dense covariance is not implemented on "
                                    "the device")
    walk.update(kind='learned', sidx=[int(i) for i in np.asarray(model._srw_idxs).reshape(-1)])
    return walk, len(walk['idxs']), 1


# ---------------------------------------------------------------------------
# GPU-backed model classes with the reference's interface
# ---------------------------------------------------------------------------

def _safe_shape(a):
    a = np.asarray(a)
    return a.shape[0] if a.ndim else 1


class Model(object):
    """The members of qinfer.abstract_model.Model the hot path consumes
    (abstract_model.py:96-117, 274-281, 311-318, 357-395, 430-468)."""

    def __init__(self):
        self._call_count = 0

    @property
    def call_count(self):
        return self._call_count

    @property
    def is_n_outcomes_constant(self):
        return True

    def n_outcomes(self, expparams):
        return 2

    @property
    def Q(self):
        """Diagonal of the quadratic-loss scale matrix (abstract_model.py:170-186): the identity by default."""
        return np.ones((self.n_modelparams,))

    def domain(self, expparams):
        """abstract_model.py:287-298: one ``IntegerDomain(0, n_outcomes - 1)`` per experiment."""
        if expparams is None:
            return IntegerDomain(min=0, max=1)
        n_o = np.broadcast_to(np.asarray(self.n_outcomes(expparams)), (_safe_shape(expparams),))
        return [IntegerDomain(min=0, max=int(k) - 1) for k in n_o]

    def clear_cache(self):
        pass

    def canonicalize(self, modelparams):
        return modelparams

    def update_timestep(self, modelparams, expparams):
        # abstract_model.py:357-374: the identity, as an (n_models, n_modelparams, n_experiments) copy
        return np.tile(modelparams, (expparams.shape[0], 1, 1)).transpose((1, 2, 0))

    def are_models_valid(self, modelparams):
        from .engine import host_are_models_valid
        return host_are_models_valid(describe_model(self), modelparams)

    def likelihood(self, outcomes, modelparams, expparams):
        """(n_outcomes, n_models, n_expparams) likelihood tensor, evaluated by the
        CUDA kernels (abstract_model.py:443-468 contract, incl. ``call_count``)."""
        from .engine import host_likelihood
        self._call_count += _safe_shape(outcomes) * _safe_shape(modelparams) * _safe_shape(expparams)
        return host_likelihood(describe_model(self), outcomes, modelparams, expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        """abstract_model.py:632-658 for constant two-outcome domains (host RNG, like the reference)."""
        all_outcomes = np.arange(2)
        probabilities = self.likelihood(all_outcomes, modelparams, expparams)
        cdf = np.cumsum(probabilities, axis=0)
        randnum = np.random.random((repeat, 1, modelparams.shape[0], expparams.shape[0]))
        outcomes = all_outcomes[np.argmax(cdf > randnum, axis=1)]
        if repeat == 1 and expparams.shape[0] == 1 and modelparams.shape[0] == 1:
            return outcomes[0, 0, 0]
        return outcomes


class IntegerDomain(object):
    """domains.py:427-560, the members the updater reads: ``values``, ``n_members``, ``min``, ``max``."""

    def __init__(self, min=0, max=1):
        self._min, self._max = int(min), int(max)

    min = property(lambda self: self._min)
    max = property(lambda self: self._max)
    n_members = property(lambda self: self._max - self._min + 1)
    is_finite = True
    dtype = int

    @property
    def values(self):
        return np.arange(self._min, self._max + 1, dtype=int)


class CoinModel(Model):
    """test_models.py:262-326: the model parameter is the probability of outcome 0; no experiment parameters."""
    n_modelparams = 1
    modelparam_names = [r'p']
    expparams_dtype = []


class SimpleInversionModel(Model):
    """test_models.py:64-166: pr0 = cos^2(t (omega - w_) / 2); valid iff omega > min_freq."""

    def __init__(self, min_freq=0):
        super(SimpleInversionModel, self).__init__()
        self._min_freq = min_freq

    n_modelparams = 1
    modelparam_names = [r'\omega']
    expparams_dtype = [('t', 'float'), ('w_', 'float')]


class SimplePrecessionModel(SimpleInversionModel):
    """test_models.py:169-213: scalar experiment parameter t, w_ = 0."""
    expparams_dtype = 'float'


class RandomizedBenchmarkingModel(Model):
    """rb.py:86-195 (zeroth order): pr0 = 1 - (A p^m + B)."""

    def __init__(self, interleaved=False, order=0):
        if order != 0:
            raise NotImplementedError("Only zeroth-order is currently implemented.")
        super(RandomizedBenchmarkingModel, self).__init__()
        self._il = interleaved

    @property
    def n_modelparams(self):
        return 3 + (1 if self._il else 0)

    @property
    def modelparam_names(self):
        return [r'\tilde{p}', 'p', 'A', 'B'] if self._il else ['p', 'A', 'B']

    @property
    def expparams_dtype(self):
        return [('m', 'uint')] + ([('reference', bool)] if self._il else [])


class MLEModel(Model):
    """derived_models.py:681-703: amplifies the Bayes update by raising every likelihood of the underlying model to
    ``likelihood_power`` (the fictional posterior of [JDD08] whose mean approximates the MLE)."""

    def __init__(self, underlying_model, likelihood_power):
        super(MLEModel, self).__init__()
        self._underlying_model = underlying_model
        self._pow = likelihood_power

    underlying_model = property(lambda self: self._underlying_model)
    decorated_model = property(lambda self: self._underlying_model)
    base_model = property(lambda self: getattr(self._underlying_model, 'base_model', self._underlying_model))
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    modelparam_names = property(lambda self: self._underlying_model.modelparam_names)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)
    is_n_outcomes_constant = property(lambda self: self._underlying_model.is_n_outcomes_constant)
    Q = property(lambda self: self._underlying_model.Q)

    def n_outcomes(self, expparams):
        return self._underlying_model.n_outcomes(expparams)

    def domain(self, expparams):
        return self._underlying_model.domain(expparams)

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def canonicalize(self, modelparams):
        return self._underlying_model.canonicalize(modelparams)

    def clear_cache(self):
        self._underlying_model.clear_cache()

    def update_timestep(self, modelparams, expparams):
        return self._underlying_model.update_timestep(modelparams, expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        return self._underlying_model.simulate_experiment(modelparams, expparams, repeat)


class _Decorator(Model):
    """derived_models.py:64-146 (DerivedModel): everything but the overridden members forwards to the decorated model."""

    def __init__(self, underlying_model):
        super(_Decorator, self).__init__()
        self._underlying_model = underlying_model

    underlying_model = property(lambda self: self._underlying_model)
    decorated_model = property(lambda self: self._underlying_model)
    base_model = property(lambda self: getattr(self._underlying_model, 'base_model', self._underlying_model))
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    modelparam_names = property(lambda self: self._underlying_model.modelparam_names)
    expparams_dtype = property(lambda self: self._underlying_model.expparams_dtype)
    is_n_outcomes_constant = property(lambda self: self._underlying_model.is_n_outcomes_constant)
    Q = property(lambda self: self._underlying_model.Q)

    def n_outcomes(self, expparams):
        return self._underlying_model.n_outcomes(expparams)

    def domain(self, expparams):
        return self._underlying_model.domain(expparams)

    def are_models_valid(self, modelparams):
        return self._underlying_model.are_models_valid(modelparams)

    def canonicalize(self, modelparams):
        return self._underlying_model.canonicalize(modelparams)

    def clear_cache(self):
        self._underlying_model.clear_cache()

    def update_timestep(self, modelparams, expparams):
        return self._underlying_model.update_timestep(modelparams, expparams)

    def likelihood(self, outcomes, modelparams, expparams):
        self._call_count += _safe_shape(outcomes) * _safe_shape(modelparams) * _safe_shape(expparams)
        return self._underlying_model.likelihood(outcomes, modelparams, expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        return self._underlying_model.simulate_experiment(modelparams, expparams, repeat)


def binom_est_error(p, N, hedge=float(0)):
    """utils.py:683-688: standard error of a (hedged) binomial estimator."""
    return np.sqrt(p * (1 - p) / (N + 2 * hedge + 1))


class PoisonedModel(_Decorator):
    """derived_models.py:148-220: the underlying likelihood plus clipped Gaussian noise mimicking the sampling error of
    adaptive (ALE: fixed tolerance ``tol``) or fixed-sample (MLE: ``n_samples``, ``hedge``) likelihood estimation.
    Inside an ``SMCUpdater`` the noise is applied on the device (qb_poison_likelihood); this host-array method follows
    the reference line by line."""

    def __init__(self, underlying_model, tol=None, n_samples=None, hedge=None):
        super(PoisonedModel, self).__init__(underlying_model)
        if (tol is None) == (n_samples is None):
            raise ValueError("Exactly one of tol and n_samples must be specified")
        self._tol = tol
        self._n_samples = n_samples
        self._hedge = (hedge if hedge is not None else 0.0) if tol is None else None

    def likelihood(self, outcomes, modelparams, expparams):
        L = np.array(super(PoisonedModel, self).likelihood(outcomes, modelparams, expparams))
        epsilon = np.random.normal(size=L.shape)
        if self._tol is not None:
            epsilon *= self._tol
        else:
            epsilon *= binom_est_error(p=L, N=self._n_samples, hedge=self._hedge)
        np.clip(L + epsilon, 0, 1, out=L)
        return L


class RandomWalkModel(_Decorator):
    """derived_models.py:705-741: after every update each particle takes a step drawn from ``step_distribution``
    (any object with ``n_rvs`` and ``sample(n)``).  Inside an ``SMCUpdater`` the steps are added on the device
    (qb_walk_step)."""

    def __init__(self, underlying_model, step_distribution):
        super(RandomWalkModel, self).__init__(underlying_model)
        self._step_dist = step_distribution
        if underlying_model.n_modelparams != step_distribution.n_rvs:
            raise TypeError("Step distribution does not match model dimension.")

    def update_timestep(self, modelparams, expparams):
        steps = self._step_dist.sample(n=modelparams.shape[0] * expparams.shape[0])
        steps = steps.reshape((modelparams.shape[0], expparams.shape[0], self.n_modelparams))
        return modelparams[:, :, np.newaxis] + steps.transpose((0, 2, 1))


class GaussianRandomWalkModel(_Decorator):
    """derived_models.py:743-963: zero-mean Gaussian steps on the parameters ``random_walk_idxs`` after every update;
    the covariance is fixed (``fixed_covariance``: its diagonal, or the dense matrix with ``diagonal=False``) or, for
    the diagonal case, learned — one extra model parameter sigma >= 0 per walking parameter.  A
    ``model_transformation`` and a learned dense covariance are not implemented on the device."""

    def __init__(self, underlying_model, random_walk_idxs='all', fixed_covariance=None, diagonal=True,
                 scale_mult=None, model_transformation=None):
        super(GaussianRandomWalkModel, self).__init__(underlying_model)
        n_u = underlying_model.n_modelparams
        self._diagonal = diagonal
        self._rw_idxs = np.s_[:n_u] if isinstance(random_walk_idxs, str) and random_walk_idxs == 'all' \
            else random_walk_idxs
        explicit = np.arange(n_u)[self._rw_idxs]
        if explicit.size == 0:
            raise IndexError('At least one model parameter must take a random walk.')
        names = list(underlying_model.modelparam_names)
        self._rw_names = [names[i] for i in explicit]
        self._n_rw = len(explicit)
        self._srw_names = []
        if fixed_covariance is None:
            self._has_fixed_covariance = False
            if not diagonal:
                raise UnsupportedModelError("GaussianRandomWalkModel with a learned dense covariance is not "
                                            "implemented on the device")
            self._srw_names = [r"\sigma_{{{}}}".format(name) for name in self._rw_names]
            self._srw_idxs = (n_u + np.arange(self._n_rw)).astype(int)
        else:
            self._has_fixed_covariance = True
            fixed_covariance = np.asarray(fixed_covariance, dtype=float)
            if diagonal:
                if fixed_covariance.ndim != 1:
                    raise ValueError('Diagonal covariance requested, but fixed_covariance has {} dimensions.'.format(
                        fixed_covariance.ndim))
                if fixed_covariance.size != self._n_rw:
                    raise ValueError('fixed_covariance dimension, {}, inconsistent with number of parameters, {}'
                                     .format(fixed_covariance.size, self._n_rw))
                self._fixed_scale = np.sqrt(fixed_covariance)
            else:
                if fixed_covariance.ndim != 2:
                    raise ValueError('Dense covariance requested, but fixed_covariance has {} dimensions.'.format(
                        fixed_covariance.ndim))
                if fixed_covariance.shape != (self._n_rw, self._n_rw):
                    raise ValueError('fixed_covariance expected to be square with width {}'.format(self._n_rw))
                self._fixed_chol = np.linalg.cholesky(fixed_covariance)
        if scale_mult is None:
            self._scale_mult_fcn = (lambda expparams: 1)
        elif isinstance(scale_mult, str):
            self._scale_mult_fcn = lambda x: x[scale_mult]
        else:
            self._scale_mult_fcn = scale_mult
        self._has_transformation = model_transformation is not None
        if self._has_transformation:
            raise UnsupportedModelError("GaussianRandomWalkModel with a model_transformation is not implemented on "
                                        "the device (it is an arbitrary host callable)")

    @property
    def modelparam_names(self):
        return list(self._underlying_model.modelparam_names) + self._srw_names

    @property
    def n_modelparams(self):
        return len(self.modelparam_names)

    @property
    def is_n_outcomes_constant(self):
        return False

    def are_models_valid(self, modelparams):
        from .engine import host_are_models_valid
        return host_are_models_valid(describe_model(self), modelparams)

    def likelihood(self, outcomes, modelparams, expparams):
        self._call_count += _safe_shape(outcomes) * _safe_shape(modelparams) * _safe_shape(expparams)
        return self._underlying_model.likelihood(
            outcomes, np.asarray(modelparams)[..., :self._underlying_model.n_modelparams], expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        return self._underlying_model.simulate_experiment(
            np.asarray(modelparams)[..., :self._underlying_model.n_modelparams], expparams, repeat)

    def update_timestep(self, modelparams, expparams):
        # derived_models.py:921-963, host arrays (the updater applies the same arithmetic on the device)
        n_mps, n_eps = modelparams.shape[0], expparams.shape[0]
        if self._diagonal:
            scale = self._fixed_scale if self._has_fixed_covariance else modelparams[:, self._srw_idxs]
            steps = scale * np.random.normal(size=(n_eps, n_mps, self._n_rw))
            steps = steps.transpose((1, 2, 0))
        else:
            steps = np.dot(self._fixed_chol, np.random.normal(size=(self._n_rw, n_mps * n_eps))
                           ).reshape(self._n_rw, n_mps, n_eps).transpose((1, 0, 2))
        steps = self._scale_mult_fcn(expparams) * steps
        new_mps = np.repeat(modelparams[:, :, np.newaxis], n_eps, axis=2)
        new_mps[:, self._rw_idxs, :] += steps
        return new_mps


class BinomialModel(Model):
    """derived_models.py:222-360: n_meas iid shots of a two-outcome model; the
    datum is the number of '1' outcomes."""

    def __init__(self, underlying_model):
        super(BinomialModel, self).__init__()
        if not (underlying_model.is_n_outcomes_constant and underlying_model.n_outcomes(None) == 2):
            raise ValueError("Decorated model must be a two-outcome model.")
        self._underlying_model = underlying_model
        if isinstance(underlying_model.expparams_dtype, str):
            self._expparams_scalar = True
            self._expparams_dtype = [('x', underlying_model.expparams_dtype), ('n_meas', 'uint')]
        else:
            self._expparams_scalar = False
            self._expparams_dtype = underlying_model.expparams_dtype + [('n_meas', 'uint')]

    underlying_model = property(lambda self: self._underlying_model)
    decorated_model = property(lambda self: self._underlying_model)
    base_model = property(lambda self: getattr(self._underlying_model, 'base_model', self._underlying_model))
    n_modelparams = property(lambda self: self._underlying_model.n_modelparams)
    modelparam_names = property(lambda self: self._underlying_model.modelparam_names)
    expparams_dtype = property(lambda self: self._expparams_dtype)

    @property
    def is_n_outcomes_constant(self):
        return False

    def n_outcomes(self, expparams):
        return expparams['n_meas'] + 1

    def canonicalize(self, modelparams):
        return self._underlying_model.canonicalize(modelparams)

    def clear_cache(self):
        self._underlying_model.clear_cache()

    def update_timestep(self, modelparams, expparams):
        return self._underlying_model.update_timestep(
            modelparams, expparams['x'] if self._expparams_scalar else expparams)

    def simulate_experiment(self, modelparams, expparams, repeat=1):
        # derived_models.py:331-355 (host binomial sampler, as in the reference)
        import scipy.stats
        pr1 = self._underlying_model.likelihood(
            np.array([1], dtype='uint'), modelparams,
            expparams['x'] if self._expparams_scalar else expparams)
        dist = scipy.stats.binom(expparams['n_meas'].astype('int'), pr1[0, :, :])
        if pr1.size != 1:
            os_ = np.concatenate([dist.rvs()[np.newaxis, :, :] for _ in range(repeat)], axis=0)
        else:
            os_ = np.concatenate([np.array([[[dist.rvs()]]]) for _ in range(repeat)], axis=0)
        return os_[0, 0, 0] if os_.size == 1 else os_


# ---- tomography ---------------------------------------------------------------

def gell_mann_basis_data(dim):
    """Generalised Gell-Mann matrices, shape (dim^2, dim, dim) (tomography/bases.py:71-111)."""
    B = np.zeros((dim ** 2, dim, dim), dtype=complex)
    B[0] = np.eye(dim) / np.sqrt(dim)
    for k in range(1, dim):
        B[k] = np.diag(np.concatenate([np.ones((k,)), [-k], np.zeros((dim - k - 1,))])) / np.sqrt(k + k ** 2)
    off = dim * (dim - 1) // 2
    for i in range(1, dim):
        for j in range(i):
            k = (i - 1) * i // 2 + j + dim
            B[k, [i, j], [j, i]] = 1 / np.sqrt(2)
            B[k + off, [i, j], [j, i]] = [1j / np.sqrt(2), -1j / np.sqrt(2)]
    return B


class TomographyBasis(object):
    """tomography/bases.py:180-321: ``data`` (dim^2, dim, dim), ``dim``, ``flat()``."""

    def __init__(self, data, dims=None, labels=None, name=None):
        self.data = np.asarray(data, dtype=complex)
        self.dims = dims if dims is not None else [self.data.shape[1]]
        self.labels = labels
        self._name = name
        self._flat = self.data.reshape((self.data.shape[0], -1))

    @property
    def dim(self):
        return int(np.prod(self.dims))

    def flat(self):
        return self._flat

    def __len__(self):
        return self.dim ** 2


def gell_mann_basis(dim):
    return TomographyBasis(gell_mann_basis_data(dim), [dim], name='gell_mann_basis')


def pauli_basis(nq=1):
    """nq-qubit Pauli basis {1, X, Y, Z}^{(x) nq} / sqrt(2^nq) (tomography/bases.py:137-153)."""
    single = gell_mann_basis_data(2)[[0, 2, 3, 1]]
    dim = 2 ** nq
    data = np.zeros((dim ** 2, dim, dim), dtype=complex)
    for idx, factors in enumerate(itertools.product(*([single] * nq))):
        data[idx] = reduce(np.kron, factors)
    return TomographyBasis(data, [2] * nq, name='pauli_basis')


class TomographyModel(Model):
    """tomography/models.py:82-226: two-outcome POVM tomography of a dim-level state."""

    def __init__(self, basis, allow_subnormalized=False):
        super(TomographyModel, self).__init__()
        self._dim = basis.dim
        self._basis = basis
        self._allow_subnormalied = allow_subnormalized

    dim = property(lambda self: self._dim)
    basis = property(lambda self: self._basis)

    @property
    def n_modelparams(self):
        return self._dim ** 2

    @property
    def expparams_dtype(self):
        return [('meas', float, self._dim ** 2)]

    def canonicalize(self, modelparams):
        from .engine import host_canonicalize
        return host_canonicalize(describe_model(self), modelparams)


class DiffusiveTomographyModel(TomographyModel):
    """tomography/models.py:228-272: tomography of a state that diffuses between measurements.  One extra model
    parameter eps > 0 (the diffusion scale per unit sqrt-time) and one extra experiment field ``t``; after every
    update each state parameter but the first takes a N(0, (eps sqrt(t))^2) step and the state is re-canonicalised."""

    @property
    def n_modelparams(self):
        return self._dim ** 2 + 1

    @property
    def expparams_dtype(self):
        return [('meas', float, self._dim ** 2), ('t', float)]

    def canonicalize(self, modelparams):
        from .engine import host_canonicalize
        modelparams = np.asarray(modelparams, dtype=float)
        return host_canonicalize(describe_model(self), modelparams)

    def update_timestep(self, modelparams, expparams):
        eps = (modelparams[:, -1, None] * np.sqrt(expparams['t']))[:, :, None]
        steps = eps * np.random.randn(*modelparams[:, None, :].shape)
        steps[:, :, [0, -1]] = 0
        raw = modelparams[:, None, :] + steps
        for idx_experiment in range(len(expparams)):
            raw[:, idx_experiment, :] = self.canonicalize(raw[:, idx_experiment, :])
        return raw.transpose((0, 2, 1))
