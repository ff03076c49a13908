"""Resampler plugin surface: ``Resampler`` ABC and the B200 ``LiuWestResampler``.

Same call contract as qinfer.resamplers (resamplers.py:73-95, 223-392):
``resampler(model, particle_dist, n_particles=None, precomputed_mean=None,
precomputed_cov=None) -> particle distribution``.  The Liu-West steps run as
CUDA kernels (moments, CDF scan, bisection draw, shrink+perturb+validity,
ordered compaction, retry); the host keeps only what the reference keeps in
scalar Python: the d x d ``sqrtm_psd`` and the control flow of the retry loop.

Two extra, keyword-only knobs that the reference does not have:

``rng``   ``'numpy'`` (default) draws the uniforms and the perturbation normals
          from the legacy global ``np.random`` stream with exactly the calls,
          shapes and order of resamplers.py:319,332 and uploads them — resample
          indices are then bit-identical to the reference under the same seed.
          ``'mt19937'`` consumes the very same global ``np.random`` stream but
          generates it on the device (the generator state is read from and
          written back to ``np.random``): indices stay bit-identical, the
          normals are within 1 ulp, and a 10^7-particle resample no longer
          waits ~0.4 s for the host generator.  Needs the default kernel.
          ``'philox'`` generates both on the device (counter-based, seeded by
          ``seed``) for throughput.
``scan``  ``'exact'`` (default) reproduces ``np.cumsum``'s sequential fp64
          rounding bit for bit; ``'fast'`` is a re-associated parallel scan.
``draw``  device-RNG mode with ``scan='fast'`` and d <= 4 only.  ``'binned'``: the multinomial draw factorised into
          counts per bin of 2048 particles + i.i.d. draws inside each bin from a shared-memory CDF; moments, counts,
          draw, move and the new weights in three streaming launches (csrc/qb_binned.cu); same law as the
          reference's draw, offspring grouped by parent bin, a retry re-centres on the particle's own parent.
          ``'guided'``: i.i.d. order through the guide table,
          bit-identical to the staged launches (incl. the reference's prefix-``mus`` retry quirk).  ``'merge'``: the
          uniforms are generated already sorted (exponential spacings, a scan), so the draw streams the CDF and
          the parents once instead of bisecting at random; the new particles come out ordered by parent and a
          retry re-centres on the particle's own parent.  ``'auto'`` (default): binned.
"""
import abc
import os
import warnings

import numpy as np
import scipy.linalg
import torch

from . import _lib
from ._exceptions import ApproximationWarning, ResamplerError, ResamplerWarning
from .distributions import ParticleDistribution, covariance_from_moments


MERGE_DRAW_MIN_PARTICLES = 32 * 1000 * 1000


def sqrtm_psd(A, est_error=True, check_finite=True):
    """PSD matrix square root through ``scipy.linalg.eigh`` with non-positive
    eigenvalues clipped — the same host call as utils.py:593-607 (d x d, d <= 64)."""
    A = np.asarray(A)
    if A.shape == (1, 1) and np.isrealobj(A):
        # eigh of a 1 x 1 matrix is (A[0,0], [[+-1]]): the same numbers without the LAPACK round trip
        if check_finite and not np.isfinite(A[0, 0]):
            raise ValueError("array must not contain infs or NaNs")
        root = np.sqrt(A[0, 0]) if A[0, 0] > 0 else 0.0
        A_sqrt = np.array([[root]], dtype=np.float64)
        if est_error:
            return A_sqrt, np.linalg.norm(np.dot(A_sqrt, A_sqrt) - A, 'fro')
        return A_sqrt
    w, v = scipy.linalg.eigh(A, check_finite=check_finite)
    w[w <= 0] = 0
    np.sqrt(w, out=w)
    A_sqrt = (v * w).dot(v.conj().T)
    if est_error:
        return A_sqrt, np.linalg.norm(np.dot(A_sqrt, A_sqrt) - A, 'fro')
    return A_sqrt


def _cov_1x1(mean, second_moment):
    """``covariance_from_moments`` for one model parameter: the same product, difference and tests as
    distributions.py:388-397 without the array machinery (the eigenvalue of a 1 x 1 matrix is its entry)."""
    c = second_moment[0, 0] - mean[0] * mean[0]
    assert np.isfinite(c)
    if not c >= 0:
        warnings.warn('Numerical error in covariance estimation causing positive semidefinite violation.',
                      ApproximationWarning)
    return np.array([[c]])


class Resampler(abc.ABC):
    @abc.abstractmethod
    def __call__(self, model, particle_dist, n_particles=None, precomputed_mean=None, precomputed_cov=None):
        """Resample ``particle_dist`` into ``n_particles`` new particles."""


class DeviceParticles(object):
    """What the device path of a resampler returns: the new particles are already
    in the cloud's alternate slab; host arrays are produced only on request."""

    def __init__(self, cloud, n_particles, weights_fused=False):
        self.cloud = cloud
        self.n_particles = int(n_particles)
        self.weights_fused = bool(weights_fused)   # the alternate weight buffer already holds 1/n and its stats

    @property
    def particle_weights(self):
        return np.ones((self.n_particles,)) / self.n_particles

    @property
    def particle_locations(self):
        return self.cloud.x_alt.cpu().numpy()

    @property
    def n_rvs(self):
        return self.cloud.d


class LiuWestResampler(Resampler):
    def __init__(self, a=0.98, h=None, maxiter=1000, debug=False, postselect=True, zero_cov_comp=1e-10,
                 default_n_particles=None, kernel=np.random.randn, *, rng='numpy', seed=None, scan='exact',
                 draw=None):
        self._default_n_particles = default_n_particles
        self._override_h = False
        self.a = a
        if h is not None:
            self._override_h = True
            self._h = h
        self._maxiter = maxiter
        self._debug = debug
        self._postselect = postselect
        self._zero_cov_comp = zero_cov_comp
        self._kernel = kernel
        if rng not in ('numpy', 'mt19937', 'philox'):
            raise ValueError("rng must be 'numpy', 'mt19937' or 'philox'")
        if scan not in ('exact', 'fast'):
            raise ValueError("scan must be 'exact' or 'fast'")
        if rng != 'numpy' and kernel is not np.random.randn:
            raise ValueError("a custom perturbation kernel needs rng='numpy' (it is a host callable)")
        if draw is None:
            draw = os.environ.get("QB_DRAW", "auto")        # (environment override: A/B runs of the bench)
        if draw not in ('auto', 'binned', 'merge', 'guided'):
            raise ValueError("draw must be 'auto', 'binned', 'merge' or 'guided'")
        self._rng = rng
        self._scan = scan
        self._draw = draw       # device-RNG mode only: sorted uniforms + streaming merge, or i.i.d. order + guide table
        self._seed = int(seed) if seed is not None else 0x5EED
        self._philox_offset = 0
        self.last_n_iters = 0
        self.last_overflow = 0
        self._fused = True      # device-RNG mode: fused draw+move kernels (False: the staged launches, same result)
        # binned draw, postselection retry: False (default) re-centres a still-invalid particle on an i.i.d. draw from
        # the weighted cloud — the LAW of the reference's loop, whose `mus = mus[:k]` (resamplers.py:372) pairs the
        # r-th invalid particle with the r-th original draw; True keeps the particle's own parent (textbook Liu-West)
        self._own_mean = False
        self._small = True      # parity mode: clouds of <= QB_SMALL_MAX particles resample in one single-CTA launch

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

    # -- helpers ---------------------------------------------------------------
    def _uniforms(self, cloud, n):
        if self._rng == 'numpy':
            u = np.random.random((n,))                       # resamplers.py:319
            cloud._u.copy_(torch.from_numpy(u))
        elif self._rng == 'mt19937':
            cloud.mt19937_uniform(cloud._u, n)
        else:
            cloud.rng_uniform(cloud._u, n, self._seed, self._philox_offset)
            self._philox_offset += (n + 1) // 2
        return cloud._u

    def _normals(self, cloud, d, k):
        buf = cloud._eps[:d * k]
        if self._rng == 'numpy':
            eps = np.ascontiguousarray(self._kernel(d, k), dtype=np.float64)   # resamplers.py:332, shape (d, k)
            if eps.shape != (d, k):
                raise ValueError("resampling kernel returned shape %s, expected %s" % (eps.shape, (d, k)))
            buf.copy_(torch.from_numpy(eps.reshape(-1)))
        elif self._rng == 'mt19937':
            cloud.mt19937_normal(buf, d * k)                 # randn(d, k) fills row-major from the same stream
        else:
            cloud.rng_normal(buf, d * k, self._seed ^ 0x9E3779B97F4A7C15, self._philox_offset)
            self._philox_offset += (d * k + 1) // 2
        return buf

    def _staged_pass(self, cloud, mean, S, a, n_particles):
        """CDF, uniforms, draw, normals, move as separate launches (any RNG, any d); u, js and eps live in HBM."""
        d = cloud.d
        cloud._resample_scratch(n_particles)
        cloud.cdf(_lib.QB_SCAN_EXACT if self._scan == 'exact' else _lib.QB_SCAN_FAST)
        cloud.draw(self._uniforms(cloud, n_particles), n_particles)
        n_iters = 0
        n_invalid = n_particles
        first = True
        while n_invalid and n_iters < self._maxiter:
            n_iters += 1
            if first:
                eps = self._normals(cloud, d, n_particles)
                cloud.lw_move(mean, S, a, eps, n_particles, self._postselect)
                first = False
            else:
                cloud.compact_invalid(n_particles)
                eps = self._normals(cloud, d, n_invalid)
                cloud.lw_retry(mean, S, a, eps, n_invalid)
            n_invalid, overflow = cloud.read_counter()
            if n_iters == 1:
                self.last_overflow = overflow
        return n_iters, n_invalid

    def _fused_pass(self, cloud, mean, S, a, n_particles, dst=None, scale_u=False, own_mean=False, seed=None,
                    build_cdf=True, auto_merge=True):
        """Device-RNG mode, d <= 4: the CDF pass also scatters the draw's guide table, and ONE kernel draws, gathers,
        shrinks, perturbs and tests validity (u, js, eps never touch HBM).  Consumes the Philox streams exactly like
        ``_staged_pass`` (uniforms, then normals per iteration), so both give bit-identical particles."""
        d = cloud.d
        seed = self._seed if seed is None else int(seed)
        seed_u, seed_n = seed, seed ^ 0x9E3779B97F4A7C15
        if build_cdf:
            cloud.cdf(_lib.QB_SCAN_FAST_GUIDE_SCALED if scale_u else _lib.QB_SCAN_FAST_GUIDE)
        # measured (B200, d = 1): guided 310 us vs merge 331 us at n = 1e7, 6.9 ms vs 3.1 ms at n = 1e8 (the guided
        # draw's random sectors run out of TLB reach / L2 there); 'auto' switches in between
        merge = self._draw == 'merge' or (self._draw == 'auto' and auto_merge and cloud.n >= MERGE_DRAW_MIN_PARTICLES)
        off_u = self._philox_offset
        self._philox_offset += (n_particles + 2) // 2 if merge else (n_particles + 1) // 2
        off_n = self._philox_offset
        self._philox_offset += (d * n_particles + 1) // 2
        if merge:
            # sorted uniforms (n + 1 exponential spacings of the uniform stream) + streaming merge with the CDF
            cloud.lw_merge_move(mean, S, a, seed_u, off_u, seed_n, off_n, n_particles, self._postselect, dst=dst,
                                scale_u=scale_u)
        else:
            cloud.lw_draw_move(mean, S, a, seed_u, off_u, seed_n, off_n, n_particles, self._postselect, dst=dst,
                               scale_u=scale_u)
        n_invalid, self.last_overflow = cloud.read_counter()
        n_iters = 1
        while n_invalid and n_iters < self._maxiter:
            n_iters += 1
            cloud.compact_invalid(n_particles)
            off_n = self._philox_offset
            self._philox_offset += (d * n_invalid + 1) // 2
            if merge:
                cloud.lw_merge_retry(mean, S, a, seed_n, off_n, n_invalid, dst=dst)
            else:
                cloud.lw_draw_retry(mean, S, a, seed_u, off_u, seed_n, off_n, n_invalid, dst=dst, scale_u=scale_u,
                                    own_mean=own_mean)
            n_invalid, _ = cloud.read_counter()
        return n_iters, n_invalid

    def _binned_offsets(self, n_particles):
        """Philox stream positions of one binned resample: bin-locating uniforms, in-bin uniforms, then normals."""
        off_u = self._philox_offset
        off_v = off_u + (n_particles + 1) // 2
        self._philox_offset = off_v + (n_particles + 1) // 2
        return off_u, off_v

    RETRY_ROUNDS = 8        # perturbations per invalid particle and retry launch (binned draw)

    def _binned_plan(self, d, n_particles):
        """Philox positions of the move's normals and of the queued retry rounds; advances the stream."""
        stride = (d * n_particles + 1) // 2
        off_n = self._philox_offset
        self._philox_offset += stride
        rounds = 0
        if self._postselect and self._maxiter > 1:
            rounds = max(1, min(self.RETRY_ROUNDS, self._maxiter - 1))
            self._philox_offset += rounds * stride
        return stride, off_n, rounds

    def _binned_finish(self, cloud, tag, rounds, stride, mean, S, a, n_particles, seed_n, dst=None, split=None,
                       dst2=None, seed_v=0):
        """Wait for the move (and the retry launch queued behind it), then run the rest of the postselection loop of
        resamplers.py:327-372 if anything is still invalid (rare: 8 rounds have already run on the device)."""
        if rounds:
            n_invalid, used, listed = cloud.binned_retry_wait(tag, queued=True)
            _, self.last_overflow, drawn = cloud.binned_counters_wait(tag)
            n_iters = 1 + (used if listed else 0)
        else:
            n_invalid, self.last_overflow, drawn = cloud.binned_counters_wait(tag)
            n_iters = 1
            if not self._postselect:
                n_invalid = 0
        if drawn != n_particles:
            raise _lib.QbError("binned resample drew %d offspring, expected %d" % (drawn, n_particles))
        while n_invalid and n_iters < self._maxiter:
            more = min(self.RETRY_ROUNDS, self._maxiter - n_iters)
            off_n = self._philox_offset
            self._philox_offset += more * stride
            t2 = cloud.binned_retry(mean, S, a, seed_n, off_n, n_particles, more, dst=dst, split=split, dst2=dst2,
                                    own_mean=self._own_mean, seed_v=seed_v)
            n_invalid, used, _ = cloud.binned_retry_wait(t2)
            n_iters += used if n_invalid == 0 else more
        return n_iters, n_invalid

    def _binned_move(self, cloud, mean, S, a, n_particles, off_v, seed=None, fuse_weights=False, n_global=None,
                     dst=None, split=None, dst2=None):
        """Pass 3 of the binned resample with host-supplied constants (precomputed moments, sharded clouds) + the
        postselection loop.  The first retry launch is queued behind the move without a host round trip."""
        seed = self._seed if seed is None else int(seed)
        seed_n = seed ^ 0x9E3779B97F4A7C15
        stride, off_n, rounds = self._binned_plan(cloud.d, n_particles)
        tags = cloud.binned_move(mean, S, a, seed, off_v, seed_n, off_n, n_particles, self._postselect, dst=dst,
                                 split=split, dst2=dst2, fuse_weights=fuse_weights, n_global=n_global,
                                 retry_rounds=rounds, own_mean=self._own_mean)
        tag = tags[0] if rounds else tags
        return self._binned_finish(cloud, tag, rounds, stride, mean, S, a, n_particles, seed_n, dst=dst, split=split,
                                   dst2=dst2, seed_v=seed)

    def _binned_device(self, cloud, n_particles, fuse_weights):
        """The whole binned resample with NO host decision between its launches: the Liu-West constants (covariance,
        zero-norm replacement, matrix square root, shifted mean) are derived by the first kernel's last block.  The
        host reads the moments when they are published and performs the reference's CHECKS (finite assert and PSD
        warning of distributions.py:388-397, zero-norm warning of resamplers.py:288-293, ResamplerError of :296-299)
        while the device is already drawing."""
        d = cloud.d
        seed, seed_n = self._seed, self._seed ^ 0x9E3779B97F4A7C15
        off_u, off_v = self._binned_offsets(n_particles)
        stride, off_n, rounds = self._binned_plan(d, n_particles)
        tag = cloud.binned_resample(n_particles, self._a, self._h, self._zero_cov_comp, seed, off_u, off_v, seed_n,
                                    off_n, self._postselect, rounds, fuse_weights, own_mean=self._own_mean)
        _, mean, m2 = cloud.binned_moments_wait(tag)
        flag, s_err = cloud.binned_flags()
        _cov_1x1(mean, m2) if d == 1 else covariance_from_moments(mean, m2)
        if flag == 1:
            warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                          "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
        if not np.isfinite(s_err):
            raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                 "Check that n_ess is not too small.")
        return self._binned_finish(cloud, tag, rounds, stride, None, None, self._a, n_particles, seed_n, seed_v=seed)

    def _small_pass(self, cloud, n_particles, fuse_weights):
        """Parity mode on a small cloud (<= QB_SMALL_MAX particles, d <= 4): the legacy variates are drawn HERE, by
        np.random itself, in the reference's order (uniforms, resamplers.py:319, then kernel(n_rvs, n), :332), and ONE
        single-CTA launch does moments, constants, np.cumsum, searchsorted, shrink + perturb and the validity test.
        The host performs the reference's checks on the published moments; the retry iterations (rare) run through
        the staged kernels."""
        d = cloud.d
        u = np.random.random((n_particles,))
        eps = np.ascontiguousarray(self._kernel(d, n_particles), dtype=np.float64)
        if eps.shape != (d, n_particles):
            raise ValueError("resampling kernel returned shape %s, expected %s" % (eps.shape, (d, n_particles)))
        tag = cloud.small_resample(u, eps, n_particles, self._a, self._h, self._zero_cov_comp, self._postselect,
                                   fuse_weights)
        _, mean, m2 = cloud.binned_moments_wait(tag)
        flag, s_err = cloud.binned_flags()
        _cov_1x1(mean, m2) if d == 1 else covariance_from_moments(mean, m2)          # (finite assert, PSD warning)
        if flag == 1:
            warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                          "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
        if not np.isfinite(s_err):
            raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                 "Check that n_ess is not too small.")
        S = cloud.small_consts()
        n_invalid, self.last_overflow, _ = cloud.binned_counters_wait(tag)
        n_iters = 1
        if not self._postselect:
            n_invalid = 0
        while n_invalid and n_iters < self._maxiter:
            n_iters += 1
            cloud.compact_invalid(n_particles)
            k_eps = self._normals(cloud, d, n_invalid)
            cloud.lw_retry(mean, S, self._a, k_eps, n_invalid)
            n_invalid, _ = cloud.read_counter()
        return n_iters, n_invalid

    def _binned_return(self, cloud, on_device, n_particles, n_iters, n_invalid, weights_fused):
        if n_invalid:
            warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                           "iterations.").format(n_invalid, self._maxiter), ResamplerWarning)
        self.last_n_iters = n_iters
        if on_device:
            return DeviceParticles(cloud, n_particles, weights_fused)
        return ParticleDistribution(particle_locations=cloud.x_alt.cpu().numpy(),
                                    particle_weights=np.ones((n_particles,)) / n_particles)

    # -- the call ------------------------------------------------------------------
    def __call__(self, model, particle_dist, n_particles=None, precomputed_mean=None, precomputed_cov=None):
        from .engine import DeviceCloud
        from .models import describe_model

        cloud = getattr(particle_dist, '_cloud', None)
        on_device = cloud is not None
        if not on_device:                  # stand-alone use on a host-held distribution: upload it
            desc = describe_model(model)
            cloud = DeviceCloud(desc, particle_dist.n_particles)
            cloud.upload_locations(particle_dist.particle_locations)
            cloud.upload_weights(particle_dist.particle_weights)

        if n_particles is None:
            n_particles = (particle_dist.n_particles if self._default_n_particles is None
                           else self._default_n_particles)
        n_particles = int(n_particles)
        device_rng = self._rng == 'philox' and self._scan == 'fast' and self._fused
        binned = device_rng and self._draw in ('auto', 'binned') and cloud.binned_supported(n_particles)
        fused = device_rng and not binned and cloud.d <= 4 and n_particles <= cloud.n
        cdf_done = False
        small = (self._small and self._rng in ('numpy', 'mt19937') and self._scan == 'exact'
                 and precomputed_mean is None and precomputed_cov is None and cloud.small_supported(n_particles))
        if small:
            weights_fused = on_device and n_particles == cloud.n
            n_iters, n_invalid = self._small_pass(cloud, n_particles, weights_fused)
            return self._binned_return(cloud, on_device, n_particles, n_iters, n_invalid, weights_fused)
        if binned and precomputed_mean is None and precomputed_cov is None:
            weights_fused = on_device and n_particles == cloud.n
            n_iters, n_invalid = self._binned_device(cloud, n_particles, weights_fused)
            return self._binned_return(cloud, on_device, n_particles, n_iters, n_invalid, weights_fused)
        if binned:
            # precomputed moments: pass 1 + 2 are queued at once, the host takes the matrix square root of the given
            # covariance while they run
            off_u, off_v = self._binned_offsets(n_particles)
            tag = cloud.binned_prepare(n_particles, self._seed, off_u)
            if precomputed_mean is None and precomputed_cov is None:
                _, mean, m2 = cloud.binned_moments_wait(tag)
                cov = _cov_1x1(mean, m2) if cloud.d == 1 else covariance_from_moments(mean, m2)
            else:
                mean = particle_dist.est_mean() if precomputed_mean is None else precomputed_mean
                cov = particle_dist.est_covariance_mtx() if precomputed_cov is None else precomputed_cov
        elif on_device and precomputed_mean is None and precomputed_cov is None:
            cloud.moments_begin()                            # one pass gives both (resamplers.py:266-273)
            if fused:
                # the CDF (+ guide) pass does not need the moments: queue it now, so that the device is busy while
                # the host waits for the moments and takes the d x d matrix square root
                cloud.cdf(_lib.QB_SCAN_FAST_GUIDE)
                cdf_done = True
            _, mean, m2 = cloud.moments_end()
            cov = covariance_from_moments(mean, m2)
        else:
            mean = particle_dist.est_mean() if precomputed_mean is None else precomputed_mean
            cov = particle_dist.est_covariance_mtx() if precomputed_cov is None else precomputed_cov

        a, h = self._a, self._h
        if (cov[0, 0] == 0) if cov.shape == (1, 1) else (scipy.linalg.norm(cov, 'fro') == 0):
            warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                          "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
            cov = self._zero_cov_comp * np.eye(cov.shape[0])
        S, S_err = sqrtm_psd(cov)
        if not np.isfinite(S_err):
            raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                 "Check that n_ess is not too small.")
        S = np.real(h * S)

        weights_fused = False
        if binned:
            weights_fused = on_device and n_particles == cloud.n
            n_iters, n_invalid = self._binned_move(cloud, mean, S, a, n_particles, off_v, fuse_weights=weights_fused)
        elif fused:
            n_iters, n_invalid = self._fused_pass(cloud, mean, S, a, n_particles, build_cdf=not cdf_done)
        else:
            n_iters, n_invalid = self._staged_pass(cloud, mean, S, a, n_particles)
        if n_invalid:
            warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                           "iterations.").format(n_invalid, self._maxiter), ResamplerWarning)
        self.last_n_iters = n_iters

        if on_device:
            return DeviceParticles(cloud, n_particles, weights_fused)
        return ParticleDistribution(particle_locations=cloud.x_alt.cpu().numpy(),
                                    particle_weights=np.ones((n_particles,)) / n_particles)
