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
                 likelihood_power=1.0):
        self.kind = kind
        self.d = int(d)
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
                                    likelihood_power=self.likelihood_power)

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
            if meas.shape[0] != self.d:
                raise ValueError("expparams['meas'] has %d entries, model has %d parameters" % (meas.shape[0], self.d))
            for c in range(self.d):
                ep.meas[c] = meas[c]
        return ep


def _name(obj):
    return type(obj).__name__


def describe_model(model):
    """Recognise a built-in model (by exact class name, so subclasses with their
    own likelihood are NOT silently mis-evaluated) and build its descriptor."""
    binomial = False
    binomial_scalar = False
    inner = model
    power = 1.0
    if _name(model) == 'MLEModel':            # derived_models.py:681-703: L ** likelihood_power, outermost decorator
        power = float(getattr(model, '_pow'))
        model = inner = model.underlying_model
    if _name(model) == 'BinomialModel':
        binomial = True
        binomial_scalar = bool(getattr(model, '_expparams_scalar'))
        inner = model.underlying_model
    name = _name(inner)
    if name in ('SimplePrecessionModel', 'SimpleInversionModel'):
        return ModelDescriptor(_lib.QB_MODEL_PRECESSION, 1, binomial=binomial,
                               min_freq=getattr(inner, '_min_freq', 0.0),
                               scalar_expparam=(name == 'SimplePrecessionModel'),
                               binomial_scalar=binomial_scalar, likelihood_power=power)
    if name == 'RandomizedBenchmarkingModel':
        il = bool(getattr(inner, '_il', False))
        return ModelDescriptor(_lib.QB_MODEL_RB, 4 if il else 3, binomial=binomial, interleaved=il,
                               binomial_scalar=binomial_scalar, likelihood_power=power)
    if name == 'CoinModel':
        return ModelDescriptor(_lib.QB_MODEL_COIN, 1, binomial=binomial, binomial_scalar=binomial_scalar, likelihood_power=power)
    if name == 'TomographyModel':
        dim = int(getattr(inner, '_dim'))
        basis = np.ascontiguousarray(np.asarray(inner._basis.data, dtype=complex))
        if dim ** 2 > _lib.QB_MAX_D:
            raise UnsupportedModelError("TomographyModel with dim=%d exceeds QB_MAX_D" % dim)
        return ModelDescriptor(_lib.QB_MODEL_TOMOGRAPHY, dim ** 2, binomial=binomial, dim=dim, basis=basis,
                               allow_subnormalized=getattr(inner, '_allow_subnormalied', False),
                               binomial_scalar=binomial_scalar, likelihood_power=power)
    raise UnsupportedModelError(
        "%s is not one of the model families the B200 kernels implement (SimplePrecessionModel, "
        "SimpleInversionModel, RandomizedBenchmarkingModel, CoinModel, tomography.TomographyModel, optionally wrapped in "
        "BinomialModel and/or MLEModel). There is no CPU fallback." % name)


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
