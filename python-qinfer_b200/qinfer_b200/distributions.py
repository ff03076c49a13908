"""Particle container and the priors the benchmark configurations use.

``ParticleDistribution`` mirrors qinfer.distributions.ParticleDistribution
(distributions.py:261-453) for HOST-held particle sets — the object a
``Resampler`` receives or returns when it is used stand-alone — with its
reductions (n_ess, mean, covariance) evaluated by the CUDA kernels.  The priors
are host code exactly as in the reference (called once per ``reset``; SURVEY §2
item 2c) and draw from the legacy global ``np.random`` stream like it.
"""
import warnings

import numpy as np

from ._exceptions import ApproximationWarning


def covariance_from_moments(mean, second_moment):
    """cov = E[x x^T] - mu mu^T and the PSD warning of distributions.py:388-397."""
    cov = second_moment - np.dot(mean[..., np.newaxis], mean[np.newaxis, ...])
    assert np.all(np.isfinite(cov))
    if not np.all(np.linalg.eigvals(cov) >= 0):
        warnings.warn('Numerical error in covariance estimation causing positive semidefinite violation.',
                      ApproximationWarning)
    return cov


class ParticleDistribution(object):
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

    @property
    def n_particles(self):
        return self.particle_locations.shape[0]

    @property
    def n_rvs(self):
        return self.particle_locations.shape[1]

    def _device_moments(self):
        from .engine import host_moments
        return host_moments(self.particle_weights, self.particle_locations)

    @property
    def n_ess(self):
        """1 / sum_i w_i^2 (distributions.py:299-307)."""
        from .engine import host_weight_stats
        norm, sumsq = host_weight_stats(self.particle_weights)
        return 1 / sumsq

    def sample(self, n=1):
        """distributions.py:320-333 (host path for host-held particles)."""
        cdf = np.cumsum(self.particle_weights)
        return self.particle_locations[np.minimum(
            cdf.searchsorted(np.random.random((n,)), side='right'), len(cdf) - 1)]

    def est_mean(self):
        return self._device_moments()[1]

    def est_meanfn(self, fn):
        return np.einsum('i...,i...', self.particle_weights, fn(self.particle_locations))

    def est_covariance_mtx(self, corr=False):
        _, mean, m2 = self._device_moments()
        cov = covariance_from_moments(mean, m2)
        if corr:
            dstd = np.sqrt(np.diag(cov))
            cov /= np.outer(dstd, dstd)
        return cov


class UniformDistribution(object):
    """distributions.py:792-827."""

    def __init__(self, ranges=np.array([[0, 1]])):
        ranges = np.asarray(ranges, dtype=float)
        if ranges.ndim == 1:
            ranges = ranges[np.newaxis, ...]
        self._ranges = ranges
        self._n_rvs = ranges.shape[0]
        self._delta = ranges[:, 1] - ranges[:, 0]

    @property
    def n_rvs(self):
        return self._n_rvs

    def sample(self, n=1):
        z = np.random.random((n, self._n_rvs))
        return self._ranges[:, 0] + self._delta * z


class PostselectedDistribution(object):
    """distributions.py:1304-1350: redraw until ``model.are_models_valid``."""

    def __init__(self, distribution, model, maxiters=100):
        self._dist = distribution
        self._model = model
        self._maxiters = maxiters

    @property
    def n_rvs(self):
        return self._dist.n_rvs

    def sample(self, n=1):
        samples = np.empty((n, self.n_rvs))
        todo = np.arange(n)
        iters = 0
        while todo.size and iters < self._maxiters:
            samples[todo] = self._dist.sample(len(todo))
            todo = todo[np.nonzero(np.logical_not(self._model.are_models_valid(samples[todo, :])))[0]]
            iters += 1
        if todo.size:
            raise RuntimeError("Did not successfully postselect within {} iterations.".format(self._maxiters))
        return samples


class GinibreTomographyPrior(object):
    """Ginibre-ensemble prior over density matrices in basis coordinates:
    X = G1 + i G2, rho = X X^H / tr(X X^H), x_a = Re tr(B_a^H rho).  Restated
    without QuTiP from tomography/distributions.py:138-141,193-196 and
    tomography/bases.py:323-336."""

    def __init__(self, basis):
        self._basis = basis
        self._dim = basis.dim

    @property
    def n_rvs(self):
        return self._dim ** 2

    def sample(self, n=1):
        d = self._dim
        X = np.random.randn(n, d, d) + 1j * np.random.randn(n, d, d)
        rho = np.einsum('nij,nkj->nik', X, X.conj())
        rho /= np.trace(rho, axis1=1, axis2=2)[:, None, None]
        return np.real(np.einsum('aij,nij->na', self._basis.data.conj(), rho))


class MultivariateNormalDistribution(object):
    """distributions.py:882-908: N(mean, cov); ``sample`` draws ``np.random.randn(n, n_rvs)`` and maps it through
    ``scipy.linalg.sqrtm(cov)`` exactly like the reference (the step distribution of ``RandomWalkModel``)."""

    def __init__(self, mean, cov):
        import scipy.linalg as la
        self.mean = np.array(mean).flatten()
        self.cov = cov
        self.invcov = la.inv(cov)

    @property
    def n_rvs(self):
        return self.mean.shape[0]

    def sample(self, n=1):
        import scipy.linalg as la
        return np.einsum("ij,nj->ni", la.sqrtm(self.cov), np.random.randn(n, self.n_rvs)) + self.mean

    def grad_log_pdf(self, x):
        return -np.dot(self.invcov, (x - self.mean).transpose()).transpose()
