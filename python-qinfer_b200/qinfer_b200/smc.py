"""``SMCUpdater`` — the reference's updater surface (smc.py:97-551) over the B200 engine.

Constructor arguments, methods, properties, warnings and errors follow
qinfer.smc.SMCUpdater; the particle cloud lives in HBM (``engine.DeviceCloud``)
and every Bayes update is ONE fused kernel launch (likelihood x weight multiply
x normalisation sum x n_ess reduction).  ``particle_locations`` /
``particle_weights`` stay readable and assignable as NumPy arrays (lazy D2H on
read, H2D on assignment) so heuristics and user pickling keep working.

Deliberate differences from the reference, all on its non-hot paths:
  * in-place mutation of the arrays returned by ``particle_locations`` /
    ``particle_weights`` does not reach the device — assign the attribute;
  * ``track_resampling_divergence`` (an O(N^2) KL estimate) is out of scope;
  * ``update_timestep`` is elided: none of the built-in models overrides the
    identity default (abstract_model.py:357-374).
"""
import warnings

import numpy as np
import torch

from ._exceptions import ApproximationWarning, ResamplerWarning
from ._lib import QB_MAX_FUSE, QB_STAT_NORM, QB_STAT_SUMSQ, QbExpparams, nvtx_range
from .distributions import covariance_from_moments
from .engine import DeviceCloud
from .models import describe_model
from .resamplers import DeviceParticles, LiuWestResampler

_EPS = np.spacing(1)


class SMCUpdater(object):
    def __init__(self, model, n_particles, prior, resample_a=None, resampler=None, resample_thresh=0.5,
                 debug_resampling=False, track_resampling_divergence=False, zero_weight_policy='error',
                 zero_weight_thresh=None, canonicalize=True, device=None, lazy=False, fuse=None, fast_math=False):
        if track_resampling_divergence:
            raise NotImplementedError("track_resampling_divergence is outside the B200 hot path (SURVEY §2 1b)")
        self._desc = describe_model(model)        # raises UnsupportedModelError: no CPU fallback
        # fast_math=True: integer powers by squaring instead of pow / log / exp in the RB and binomial likelihoods
        # (relative deviation <= 1e-13 from the reference's operation sequence; the default keeps that sequence)
        self._desc.c_model.fast_math = 1 if fast_math else 0
        self._device = device
        self._cloud = None
        self._resample_count = 0
        self._min_n_ess = n_particles
        self.model = model
        self.prior = prior
        self._canonicalize = bool(canonicalize)
        self._debug_resampling = debug_resampling
        if resample_a is not None:
            warnings.warn("The 'resample_a' keyword argument is deprecated; use "
                          "'resampler=LiuWestResampler(a)' instead.", DeprecationWarning)
            if resampler is not None:
                raise ValueError("Both a resample_a and an explicit resampler were provided; please provide only one.")
            self.resampler = LiuWestResampler(a=resample_a)
        elif resampler is None:
            self.resampler = LiuWestResampler(default_n_particles=n_particles)
        else:
            self.resampler = resampler
        self.resample_thresh = resample_thresh
        self._just_resampled = False
        self._data_record = []
        self._normalization_record = []
        self._resampling_divergences = None
        self._zero_weight_policy = zero_weight_policy
        self._zero_weight_thresh = zero_weight_thresh if zero_weight_thresh is not None else 10 * np.spacing(1)
        self._host_locs = None
        self._host_weights = None
        self._n_ess = float(n_particles)
        # lazy=True: update() returns as soon as its kernel is queued; the bookkeeping of a step (records,
        # warnings, zero-weight policy, resample trigger) runs when the next call needs it, and the next
        # update is launched speculatively behind it (it cancels itself on the device if the step turns out
        # to need the host).  lazy=False (default) keeps the reference's call-by-call semantics exactly.
        # With lazy=True consecutive updates are additionally FUSED: up to ``fuse`` (default QB_MAX_FUSE = 8, 1 for
        # tomography) buffered updates go out as one kernel launch that reads and writes the cloud once.
        self._lazy = bool(lazy)
        # f4 decorators (RandomWalkModel, GaussianRandomWalkModel, DiffusiveTomographyModel, PoisonedModel): the
        # particles move (or the likelihood takes fresh noise) after / at every update, in the reference's order of
        # random draws -> one update per launch, settled before the next one goes out
        self._time_dependent = self._desc.walk is not None
        if self._time_dependent or self._desc.poison is not None:
            self._lazy = False
        fuse_cap = 1 if (self._desc.kind == 3 or self._desc.likelihood_power != 1.0 or self._time_dependent
                         or self._desc.poison is not None or self._desc.d_extra) else QB_MAX_FUSE
        self._fuse = fuse_cap if fuse is None else max(1, min(int(fuse), fuse_cap))
        self._settling = False
        self._counted = None        # models of the chain that keep a call count
        self._queue = []            # buffered, not yet launched: (ep_record, outcome, check_for_resample)
        self._pending = None        # launched, not yet settled: (tag, [steps])
        self.reset(n_particles)

    # ---- bookkeeping properties (smc.py:182-259) ----------------------------------
    data_record = property(lambda self: self._data_record[:])
    resampling_divergences = property(lambda self: self._resampling_divergences)

    @property
    def resample_count(self):
        self._flush()
        return self._resample_count

    @property
    def just_resampled(self):
        self._flush()
        return self._just_resampled

    @property
    def normalization_record(self):
        self._flush()
        return self._normalization_record

    @property
    def min_n_ess(self):
        self._flush()
        return self._min_n_ess

    @property
    def log_total_likelihood(self):
        return np.sum(np.log(self.normalization_record))

    # ---- particle container surface (distributions.py:289-453) ---------------------
    @property
    def n_particles(self):
        return self._cloud.n

    @property
    def n_rvs(self):
        return self._cloud.d

    @property
    def n_ess(self):
        self._flush()
        return self._n_ess

    @property
    def particle_locations(self):
        self._flush()
        if self._host_locs is None:
            self._host_locs = self._cloud.download_locations()
        return self._host_locs

    @particle_locations.setter
    def particle_locations(self, value):
        self._flush()
        value = np.asarray(value, dtype=np.float64)
        if value.shape[0] != self._cloud.n:
            self._rebuild_cloud(value.shape[0])
        self._cloud.upload_locations(value)
        self._host_locs = None

    @property
    def particle_weights(self):
        self._flush()
        if self._host_weights is None:
            self._host_weights = self._cloud.download_weights()
        return self._host_weights

    @particle_weights.setter
    def particle_weights(self, value):
        self._flush()
        value = np.asarray(value, dtype=np.float64)
        self._cloud.upload_weights(value)
        st = self._cloud.read_stats()
        self._n_ess = self._ness_from(st[QB_STAT_NORM], st[QB_STAT_SUMSQ], normalised=True)
        self._host_weights = None

    def _rebuild_cloud(self, n):
        launches = self._cloud.launches if self._cloud is not None else 0
        self._cloud = DeviceCloud(self._desc, n, self._device, capacity=self._cloud_capacity(n))
        self._cloud.launches = launches
        # the decorators' noise follows the resampler's generator choice: 'numpy' (default) draws from np.random on
        # the host in the reference's order, 'mt19937' continues that stream on the device, 'philox' is device-only
        res = getattr(self, 'resampler', None)
        self._cloud.noise_rng = getattr(res, '_rng', 'numpy')
        self._cloud.noise_seed = (int(getattr(res, '_seed', 0)) ^ 0x6E6F697365) & ((1 << 64) - 1)
        self._host_locs = self._host_weights = None

    def _cloud_capacity(self, n):
        """Particles to allocate for a cloud of ``n`` (a sharded cloud adds slack for its floating slabs)."""
        return n

    @staticmethod
    def _ness_from(norm, sumsq, normalised=False):
        """n_ess = 1 / sum(w_normalised^2) from the unnormalised reductions."""
        with np.errstate(divide='ignore', invalid='ignore'):
            if normalised or abs(norm) < _EPS:
                return float(np.float64(1.0) / np.float64(sumsq))
            return float((np.float64(norm) * np.float64(norm)) / np.float64(sumsq))

    def sample(self, n=1):
        """distributions.py:320-333: weighted draw (clamped), host RNG like the reference."""
        from . import _lib
        self._flush()
        cloud = self._cloud
        cloud._resample_scratch(cloud.n if cloud._js is None else cloud._js.numel())
        cloud.cdf(_lib.QB_SCAN_EXACT)
        u = torch.from_numpy(np.random.random((n,))).to(cloud.device)
        js = torch.empty((n,), dtype=torch.int64, device=cloud.device)
        from .engine import _ptr, _stream
        _lib.check(cloud.lib.qb_draw(_ptr(cloud._cdf), cloud.n, _ptr(u), n, _ptr(js), _ptr(cloud.counter[1:]),
                                     _ptr(cloud.ws), cloud.ws_bytes, _stream()))
        return cloud.x.index_select(0, js).cpu().numpy()

    def est_mean(self):
        self._flush()
        return self._cloud.moments()[1]

    def est_meanfn(self, fn):
        """distributions.py:411-430: sum_i w_i fn(x_i).  ``fn`` is first offered the DEVICE tensor of particle
        locations: a function written with arithmetic operators / torch-compatible calls (``lambda x: x ** 2``) is
        evaluated and reduced on the GPU and only its mean comes back.  A function that needs NumPy arrays raises on
        the device tensor and is applied to the downloaded cloud as in the reference."""
        self._flush()
        cloud = self._cloud
        try:
            vals = fn(cloud.x)
            on_device = isinstance(vals, torch.Tensor) and vals.is_cuda and vals.shape[0] == cloud.n
        except Exception:
            on_device = False
        if on_device:
            vals = vals.to(torch.float64)
            shape = tuple(vals.shape[1:])
            return cloud.weighted_mean_of(vals.reshape(cloud.n, -1)).reshape(shape)
        return np.einsum('i...,i...', self.particle_weights, fn(self.particle_locations))

    # ---- read-side estimators (distributions.py:457-465, 557-626) on the device: the cloud is not
    # downloaded; what travels is proportional to the answer
    def est_entropy(self):
        self._flush()
        return self._cloud.entropy()

    def est_credible_region(self, level=0.95, return_outside=False, modelparam_slice=None):
        """distributions.py:558-614: the particles of highest weight whose cumulative weight reaches ``level``, sorted
        by weight (descending).  The members are selected on the device (radix selection of the threshold weight,
        ordered compaction) and only they are downloaded; ``return_outside=True`` needs every particle and takes the
        host route."""
        self._flush()
        s_ = np.s_[modelparam_slice] if modelparam_slice is not None else np.s_[:]
        if return_outside:
            mps = self.particle_locations[:, s_]
            id_sort = np.argsort(self.particle_weights)[::-1]
            cumsum_weights = np.cumsum(self.particle_weights[id_sort])
            id_cred = cumsum_weights <= level
            id_cred[np.sum(id_cred)] = True
            return mps[id_sort][id_cred], mps[id_sort][np.logical_not(id_cred)]
        idx = self._cloud.credible_members(level)
        locs, wts = self._cloud.gather_members(idx)
        order = np.argsort(-wts, kind='stable')
        return locs[order][:, s_]

    def region_est_hull(self, level=0.95, modelparam_slice=None):
        from scipy.spatial import ConvexHull
        points = self.est_credible_region(level=level, modelparam_slice=modelparam_slice)
        hull = ConvexHull(points)
        verts = hull.vertices.flatten()
        _, first = np.unique(verts, return_index=True)          # utils.uniquify: order-preserving de-duplication
        return points[hull.simplices], points[verts[np.sort(first)]]

    def est_covariance_mtx(self, corr=False):
        self._flush()
        _, mean, m2 = self._cloud.moments()
        cov = covariance_from_moments(mean, m2)
        if corr:
            dstd = np.sqrt(np.diag(cov))
            cov /= np.outer(dstd, dstd)
        return cov

    # ---- initialisation (smc.py:281-320) ---------------------------------------------
    def reset(self, n_particles=None, only_params=None, reset_weights=True):
        if n_particles is not None and only_params is not None:
            raise ValueError("Cannot set both n_particles and only_params.")
        if self._cloud is not None:
            self._flush()
        if n_particles is None:
            n_particles = self._cloud.n
        if self._cloud is None or self._cloud.n != n_particles:
            self._rebuild_cloud(n_particles)
        cloud = self._cloud
        if reset_weights:
            cloud.set_uniform_weights()
            self._n_ess = float(n_particles)
            self._host_weights = None
        sample = np.asarray(self.prior.sample(n=n_particles), dtype=np.float64)
        if only_params is None:
            locs = sample
        else:
            locs = self.particle_locations.copy()
            locs[:, only_params] = sample[:, only_params]
        cloud.upload_locations(locs)
        self._host_locs = None
        if self._canonicalize:
            cloud.canonicalize()

    # ---- updates (smc.py:324-487) --------------------------------------------------------
    def _count_calls(self, n):
        chain = self._counted
        if chain is None:
            chain, m = [], self.model
            while m is not None:
                if hasattr(m, '_call_count'):
                    chain.append(m)
                m = getattr(m, 'underlying_model', None)
            self._counted = chain
        for m in chain:
            m._call_count += n

    @nvtx_range('qb.hypothetical_update')
    def hypothetical_update(self, outcomes, expparams, return_likelihood=False, return_normalization=False):
        """smc.py:324-386: posterior weights of hypothetical data, shape (n_outcomes, n_expparams, n_particles),
        computed on the device (likelihood, weight product, normalisation sum, division) and returned as host arrays."""
        self._flush()
        if not isinstance(outcomes, np.ndarray):
            outcomes = np.array([outcomes])
        expparams = np.atleast_1d(expparams)
        self._count_calls(outcomes.shape[0] * self._cloud.n * expparams.shape[0])
        norm_weights, L, norm_scale = self._cloud.hypothetical_update(outcomes, expparams, return_likelihood)
        out = (norm_weights,)
        if return_likelihood:
            out += (L,)
        if return_normalization:
            out += (norm_scale,)
        return out[0] if len(out) == 1 else out

    # ---- experiment design (smc.py:553-663) --------------------------------------------------
    def _design(self, expparams, want_kld):
        """Per experiment: outcome list, N (n_o,), and the device reductions of qb_design_sums."""
        self._flush()
        expparams = np.atleast_1d(expparams)
        n_eps = expparams.shape[0]
        centre = self._cloud.moments()[1]
        model = self.model
        for e in range(n_eps):
            ep1 = expparams[e:e + 1]
            if hasattr(model, 'domain'):
                os_ = np.asarray(model.domain(ep1)[0].values)
            else:
                os_ = np.arange(int(np.ravel(model.n_outcomes(ep1))[0]))
            # hypothetical_update is handed os[:-1] (smc.py:584-589): that is what the reference's call_count sees
            self._count_calls((os_.shape[0] - 1) * self._cloud.n)
            sums, kld = self._cloud.design_sums(expparams, e, os_, centre, want_kld)
            yield e, centre, sums, kld

    def bayes_risk(self, expparams):
        """smc.py:553-605: Bayes risk (quadratic loss with the model's ``Q``) of each hypothetical experiment, shape
        ``(expparams.size,)``.  The hypothetical posterior of every outcome is reduced on the device to its
        normalisation, mean and second moment about the current mean; the (n_outcomes, n_particles) weight tensor
        the reference materialises never exists."""
        expparams = np.atleast_1d(expparams)
        Q = np.asarray(getattr(self.model, 'Q', np.ones((self._cloud.d,))), dtype=np.float64)
        d = self._cloud.d
        risk = np.empty(expparams.shape[0])
        for e, centre, sums, _ in self._design(expparams, False):
            N = sums[:, 0]
            B, C = sums[:, 1:1 + d], sums[:, 1 + d:]
            with np.errstate(divide='ignore', invalid='ignore'):
                # hypothetical_update's |N| < eps guard for the outcomes it is given, plain division for the last one
                div = N.copy()
                div[:-1][np.abs(N[:-1]) < _EPS] = 1.0
                s = N / div                                       # sum of the hypothetical weights (1 up to rounding)
                delta = B / div[:, None] - centre * (1.0 - s)[:, None]    # hypothetical mean minus the centre
                var = np.sum(Q * (C / div[:, None] - 2.0 * delta * B / div[:, None]
                                  + delta ** 2 * s[:, None]), axis=1)
                risk[e] = np.sum(N * var)
        return risk

    def expected_information_gain(self, expparams):
        """smc.py:607-657: expected KL divergence of the hypothetical posterior from the current one."""
        expparams = np.atleast_1d(expparams)
        gain = np.empty(expparams.shape[0])
        for e, _, sums, kld in self._design(expparams, True):
            gain[e] = np.sum(sums[:, 0] * kld)
        return gain

    def risk(self, x0):
        return self.bayes_risk(np.array([(x0,)], dtype=self.model.expparams_dtype))

    @nvtx_range('qb.update')
    def update(self, outcome, expparams, check_for_resample=True):
        """smc.py:388-457.  With ``lazy=False`` the call returns after the step's bookkeeping exactly like the
        reference; with ``lazy=True`` it only buffers the datum (launching a fused kernel every ``fuse`` updates)."""
        self._data_record.append(outcome)
        self._enqueue(outcome, expparams, check_for_resample)
        if not self._lazy:
            self._flush()
        elif len(self._queue) >= self._fuse:
            self._launch_queue()

    def _enqueue(self, outcome, expparams, check_for_resample):
        ep = self._desc.fill_record(QbExpparams(), expparams, 0)
        if self._time_dependent:
            ep.host_expparams = np.atleast_1d(expparams)   # update_timestep's argument (scale_mult, 't')
        self._queue.append((ep, int(outcome), bool(check_for_resample)))
        self._count_calls(self._cloud.n)
        if self._pending is None and len(self._queue) == 1:
            self._just_resampled = False         # nothing in flight: new data has arrived since the last resample

    def _launch_queue(self):
        """Launch the buffered updates, at most ``fuse`` per kernel, each launch speculatively behind the pending one."""
        while self._queue:
            cloud = self._cloud
            steps, self._queue = self._queue[:self._fuse], self._queue[self._fuse:]
            prev = self._pending
            if prev is None:
                tag = cloud.fused_update(steps, cloud.cur, zero_weight_thresh=self._zero_weight_thresh,
                                         resample_below=self.n_particles * self.resample_thresh)
                self._pending = (tag, steps)
                continue
            # speculative: queue this launch behind the pending one.  It reads the pending launch's output buffers
            # and cancels itself on the device if that launch needs the host (clip, zero-weight policy, resample).
            tag = cloud.fused_update(steps, 1 - cloud.cur, guard=True, zero_weight_thresh=self._zero_weight_thresh,
                                     resample_below=self.n_particles * self.resample_thresh)
            self._queue = steps + self._queue           # provisional: they only count if `prev` was plain
            self._pending = None
            plain = self._finalize(prev)                # may warn / raise / resample, like the reference
            if plain:
                self._queue = self._queue[len(steps):]
                self._pending = (tag, steps)
            else:
                cloud._chain_tag = 0                # never chain behind a launch that cancelled itself
            # (the speculative launch cancelled itself; its steps are still queued, behind whatever steps of
            # `prev` _finalize put back)

    @nvtx_range('qb.flush')
    def _flush(self):
        """Launch and settle everything that is buffered or pending."""
        if self._settling:          # re-entered from the step being settled (resample(), est_mean(), ...)
            return
        while self._queue or self._pending is not None:
            if self._pending is not None:
                prev, self._pending = self._pending, None
                self._finalize(prev)
            if self._queue:
                self._launch_queue()

    def _finalize(self, pending):
        """The host half of smc.py:413-457 for the updates of one launch.  Returns True if every step was a plain
        commit; otherwise the state is brought to exactly what the reference holds after the step that needed the
        host (re-issuing the steps before it if necessary), the remaining steps go back to the front of the queue,
        and False is returned (a speculative successor has cancelled itself)."""
        self._settling = True
        try:
            return self._finalize_locked(pending)
        finally:
            self._settling = False

    def _finalize_locked(self, pending):
        tag, steps = pending
        cloud = self._cloud
        slot = 1 - cloud.cur
        k = len(steps)
        blocks = cloud.wait_stats(slot, tag, k)
        zt = self._zero_weight_thresh
        below = self.n_particles * self.resample_thresh
        if blocks[0][7] >= 2.0:
            # a speculative launch that cancelled itself although the host expected it to run (cannot happen while
            # host and device apply the same tests; kept as a safety net): run it again, plainly
            self._queue = steps + self._queue
            return False
        for j in range(k):
            S, Q, nbad, _, rec, ness, _, flag = blocks[j]
            ep, outcome, check = steps[j]
            self._just_resampled = False                 # smc.py:410, at the point this datum is accounted for
            if flag == 0.0:
                # plain step: the bookkeeping of smc.py:441-457, nothing else
                self._normalization_record.append(float(rec))
                self._n_ess = float(ness)
                if self._time_dependent:                 # smc.py:447-449 (a time-dependent model is never fused: k = 1)
                    cloud.commit_update()
                    self._host_weights = None
                    self._timestep(ep)
                    if self._n_ess <= self._min_n_ess:
                        self._min_n_ess = self._n_ess
                    if check and ness <= 10:
                        warnings.warn("Extremely small n_ess encountered ({}). Resampling is likely to fail. Consider "
                                      "adding particles, or resampling more often.".format(ness), ApproximationWarning)
                    return True
                if self._n_ess <= self._min_n_ess:
                    self._min_n_ess = self._n_ess
                if check and ness <= 10:
                    warnings.warn("Extremely small n_ess encountered ({}). Resampling is likely to fail. Consider "
                                  "adding particles, or resampling more often.".format(ness), ApproximationWarning)
                continue
            # ---- step j needs the host ------------------------------------------------------------------------
            self._queue = steps[j + 1:] + self._queue
            unnormalised = abs(rec) < _EPS               # smc.py:369-370: then the weights stay as w*L
            total = float(rec) if unnormalised else 1.0  # np.sum of the normalised weights
            rejected = (nbad == 0 and total <= zt and self._zero_weight_policy in ('skip', 'error'))
            keep = j if rejected else j + 1              # how many steps of this launch the cloud must reflect
            if keep < k:
                if keep > 0:
                    t2 = cloud.fused_update(steps[:keep], cloud.cur, zero_weight_thresh=zt, resample_below=below)
                    cloud.wait_stats(slot, t2, keep)
            if rejected:
                if keep > 0:
                    cloud.commit_update()
                    self._host_weights = None
                if self._zero_weight_policy == 'error':
                    raise RuntimeError("All particle weights are zero.")
                return False                             # 'skip': smc.py:427-428
            self._settle_step(slot, float(S), float(Q), float(nbad), float(rec), float(ness), check, ep)
            return False
        cloud.commit_update()                            # smc.py:441 for the whole launch
        self._host_weights = None
        return True

    def _timestep(self, ep):
        """smc.py:447-449: ``particle_locations = model.update_timestep(particle_locations, expparams)[:, :, 0]``
        for the time-dependent decorators, on the device."""
        self._cloud.walk_step(ep.host_expparams)
        self._host_locs = None

    def _settle_step(self, slot, S, Q, nbad, rec, ness, check, ep=None):
        """smc.py:416-457 for a step that needs the host; the pending buffers hold the weights after that step."""
        cloud = self._cloud
        unnormalised = abs(rec) < _EPS
        total = rec if unnormalised else 1.0
        norm_rec = rec
        if nbad > 0:                                         # smc.py:416-418
            smallest = self._pending_min_weight(slot)
            smallest = smallest if unnormalised else smallest / S
            warnings.warn("Negative weights occured in particle approximation. Smallest weight observed == {}. "
                          "Clipping weights.".format(smallest), ApproximationWarning)
            st2 = self._clip_weights(slot)
            total = float(st2[QB_STAT_NORM])
            ness = self._ness_from(total, float(st2[QB_STAT_SUMSQ]), normalised=True)
        if total <= self._zero_weight_thresh:                # smc.py:423-436 (a NaN total passes, as in the reference)
            policy = self._zero_weight_policy
            if policy == 'ignore':
                pass
            elif policy == 'warn':
                warnings.warn("All particle weights are zero. This will very likely fail quite badly.",
                              ApproximationWarning)
            elif policy == 'reset':
                warnings.warn("All particle weights are zero. Resetting from initial prior.", ApproximationWarning)
                self.reset()
                cloud = self._cloud
            elif policy in ('skip', 'error'):                # only reachable after a clip
                if policy == 'error':
                    raise RuntimeError("All particle weights are zero.")
                return
            else:
                raise ValueError("Invalid zero-weight policy {} encountered.".format(policy))
            with np.errstate(divide='ignore', invalid='ignore'):
                ness = self._ness_from(S, Q, normalised=unnormalised) if nbad == 0 else ness
        cloud.commit_update()                                # smc.py:441
        self._host_weights = None
        self._normalization_record.append(norm_rec)          # smc.py:444
        if self._time_dependent and ep is not None:          # smc.py:447-449
            self._timestep(ep)
        self._n_ess = ness
        if self._n_ess <= self._min_n_ess:                   # smc.py:452-453
            self._min_n_ess = self._n_ess
        if check:
            self._maybe_resample()

    # (hooks: a sharded cloud makes these two global with collectives)
    def _pending_min_weight(self, slot):
        return self._cloud.pending_min_weight(slot)

    def _clip_weights(self, slot):
        return self._cloud.clip_weights(slot)

    @nvtx_range('qb.batch_update')
    def batch_update(self, outcomes, expparams, resample_interval=5):
        """smc.py:459-487.  The reference loops ``update(check_for_resample=False)`` and calls ``_maybe_resample``
        after every ``resample_interval``-th datum; here the same sequence is buffered and goes out as fused
        launches (up to QB_MAX_FUSE updates each, the cloud is read and written once per launch), then settled."""
        n_exps = outcomes.shape[0]
        if expparams.shape[0] != n_exps:
            raise ValueError("The number of outcomes and experiments must match.")
        if len(expparams.shape) == 1:
            expparams = expparams[:, None]
        self._flush()
        for idx_exp, (outcome, experiment) in enumerate(zip(iter(outcomes), iter(expparams))):
            self._data_record.append(outcome)
            self._enqueue(outcome, experiment, (idx_exp + 1) % resample_interval == 0)
            if len(self._queue) >= self._fuse:
                self._launch_queue()
        self._flush()

    # ---- resampling (smc.py:263-277, 491-551) --------------------------------------------
    def _maybe_resample(self):
        self._flush()
        ess = self._n_ess
        if ess <= 10:
            warnings.warn("Extremely small n_ess encountered ({}). Resampling is likely to fail. Consider adding "
                          "particles, or resampling more often.".format(ess), ApproximationWarning)
        if ess < self.n_particles * self.resample_thresh:
            self.resample()
            return True
        return False

    @nvtx_range('qb.resample')
    def resample(self):
        self._flush()
        if self._just_resampled:
            warnings.warn("Resampling without additional data; this may not perform as desired.", ResamplerWarning)
        self._just_resampled = True
        self._resample_count += 1
        if self._debug_resampling:
            old_mean, old_cov = self.est_mean(), self.est_covariance_mtx()
        ev = None
        if self._cloud.resample_events is not None:      # bench instrumentation
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()

        new_dist = self.resampler(self.model, self)
        if isinstance(new_dist, DeviceParticles) and new_dist.cloud is self._cloud:
            self._cloud.adopt_binned(new_dist.n_particles, getattr(new_dist, 'weights_fused', False))
            # uniform weights 1/n: the stats block the kernel wrote is {norm 1, sumsq 1/n}; same two roundings here,
            # without a device read
            self._n_ess = self._ness_from(1.0, np.float64(1.0) / np.float64(new_dist.n_particles), normalised=True)
        else:                                                # a foreign resampler working on host arrays
            locs = np.asarray(new_dist.particle_locations, dtype=np.float64)
            weights = np.asarray(new_dist.particle_weights, dtype=np.float64)
            if locs.shape[0] != self._cloud.n:
                self._rebuild_cloud(locs.shape[0])
            self._cloud.upload_locations(locs)
            self._cloud.upload_weights(weights)
            st = self._cloud.read_stats()
            self._n_ess = self._ness_from(st[QB_STAT_NORM], st[QB_STAT_SUMSQ], normalised=True)
        self._host_locs = self._host_weights = None

        if self._canonicalize:
            self._cloud.canonicalize()
        try:
            self.model.clear_cache()
        except Exception as e:  # pragma: no cover
            warnings.warn("Exception raised when clearing model cache: {}. Ignoring.".format(e))
        if ev is not None:
            ev[1].record()
            self._cloud.resample_events.append(ev)

        if self._debug_resampling:
            import logging
            new_mean, new_cov = self.est_mean(), self.est_covariance_mtx()
            logging.getLogger(__name__).debug("Resampling changed mean by {}. Norm change in cov: {}.".format(
                old_mean - new_mean, np.linalg.norm(new_cov - old_cov)))
