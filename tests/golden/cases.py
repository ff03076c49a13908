"""Shared, implementation-agnostic drivers for the golden cases.

Each ``run_*`` takes a namespace ``ns`` (the reference ``qinfer`` adapter, the
oracle, or the B200 engine adapter) exposing the same plugin names, and drives
it with identical inputs.  ``make_golden.py`` runs them against the unmodified
reference to write ``*.npz``; the tests run them against the oracle (bit-exact)
and the CUDA engine (tolerances stated in the tests).
"""
import numpy as np


class Namespace(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def oracle_namespace():
    import smc_oracle as o
    return Namespace(
        name="oracle",
        SMCUpdater=o.SMCUpdater, LiuWestResampler=o.LiuWestResampler,
        ParticleDistribution=o.ParticleDistribution,
        SimplePrecessionModel=o.SimplePrecessionModel, SimpleInversionModel=o.SimpleInversionModel,
        RandomizedBenchmarkingModel=o.RandomizedBenchmarkingModel, BinomialModel=o.BinomialModel,
        CoinModel=o.CoinModel, MLEModel=o.MLEModel, TomographyModel=o.TomographyModel,
        DiffusiveTomographyModel=o.DiffusiveTomographyModel,
        RandomWalkModel=o.RandomWalkModel, GaussianRandomWalkModel=o.GaussianRandomWalkModel,
        NormalStepDistribution=o.NormalStepDistribution, PoisonedModel=o.PoisonedModel, pauli_basis=o.pauli_basis, gell_mann_basis=o.gell_mann_basis,
        UniformDistribution=o.UniformDistribution, PostselectedDistribution=o.PostselectedDistribution,
        sqrtm_psd=o.sqrtm_psd)


def reference_namespace():
    from ref_import import import_reference
    q = import_reference()
    import qinfer.tomography as qt
    import qinfer.utils as qu
    return Namespace(
        name="reference",
        SMCUpdater=q.SMCUpdater, LiuWestResampler=q.LiuWestResampler,
        ParticleDistribution=q.ParticleDistribution,
        SimplePrecessionModel=q.SimplePrecessionModel, SimpleInversionModel=q.SimpleInversionModel,
        RandomizedBenchmarkingModel=q.RandomizedBenchmarkingModel, BinomialModel=q.BinomialModel,
        CoinModel=q.CoinModel, MLEModel=q.MLEModel, TomographyModel=qt.TomographyModel,
        DiffusiveTomographyModel=qt.DiffusiveTomographyModel,
        RandomWalkModel=q.RandomWalkModel, GaussianRandomWalkModel=q.GaussianRandomWalkModel,
        NormalStepDistribution=q.MultivariateNormalDistribution, PoisonedModel=q.PoisonedModel, pauli_basis=qt.pauli_basis, gell_mann_basis=qt.gell_mann_basis,
        UniformDistribution=q.UniformDistribution, PostselectedDistribution=q.PostselectedDistribution,
        sqrtm_psd=qu.sqrtm_psd)


class FixedPrior(object):
    """A prior that hands back a pre-drawn sample (keeps prior sampling out of
    the RNG stream so every implementation starts from identical particles)."""

    def __init__(self, sample):
        self._sample = np.array(sample, dtype=float)

    @property
    def n_rvs(self):
        return self._sample.shape[1]

    def sample(self, n=1):
        assert n == self._sample.shape[0]
        return self._sample.copy()


def recording_resampler(ns, log, **kw):
    """Subclass the namespace's LiuWestResampler to snapshot its inputs/outputs."""
    Base = ns.LiuWestResampler

    class Recording(Base):
        def __call__(self, model, particle_dist, *a, **k):
            ev = dict(w=np.array(particle_dist.particle_weights, dtype=float, copy=True),
                      x=np.array(particle_dist.particle_locations, dtype=float, copy=True),
                      rng=np.random.get_state())
            out = Base.__call__(self, model, particle_dist, *a, **k)
            ev['new_x'] = np.array(out.particle_locations, dtype=float, copy=True)
            log.append(ev)
            return out

    return Recording(**kw)


def pack_rng_state(st):
    return dict(key=np.asarray(st[1], dtype=np.uint32), pos=np.int64(st[2]),
                has_gauss=np.int64(st[3]), cached=np.float64(st[4]))


def unpack_rng_state(d, prefix):
    return ('MT19937', d[prefix + 'key'], int(d[prefix + 'pos']),
            int(d[prefix + 'has_gauss']), float(d[prefix + 'cached']))


def exp_sparse_times(n, base=9.0 / 8.0):
    return base ** np.arange(n)


# ---------------------------------------------------------------------------
# Inputs (generated once, with a private RandomState — never the global stream)
# ---------------------------------------------------------------------------

def precession_inputs(n_particles=1000, n_updates=100, true_omega=0.5, seed=1234):
    rs = np.random.RandomState(seed)
    prior = rs.random_sample((n_particles, 1))
    ts = exp_sparse_times(n_updates)
    pr0 = np.cos(ts * true_omega / 2) ** 2
    outcomes = (rs.random_sample(n_updates) >= pr0).astype(np.int64)
    return dict(prior=prior, ts=ts, outcomes=outcomes)


def rb_inputs(n_particles=2000, n_updates=60, n_meas=25, seed=4321):
    rs = np.random.RandomState(seed)
    lo = np.array([0.8, 0.0, 0.0])
    hi = np.array([1.0, 1.0, 1.0])
    prior = np.empty((0, 3))
    while prior.shape[0] < n_particles:       # postselected uniform box (simple_est.py:223-236)
        c = lo + (hi - lo) * rs.random_sample((n_particles, 3))
        p, A, B = c.T
        ok = (A + B <= 1) & (A * p + B <= 1)
        prior = np.concatenate([prior, c[ok]])[:n_particles]
    p, A, B = 0.995, 0.5, 0.5
    ms = np.linspace(1, 800, n_updates).astype(int)
    counts = rs.binomial(n_meas, A * p ** ms + B)
    return dict(prior=prior, ms=ms, counts=counts.astype(np.int64), n_meas=np.int64(n_meas))


def ginibre_coords(rs, n, basis_data):
    d = basis_data.shape[1]
    X = rs.randn(n, d, d) + 1j * rs.randn(n, d, d)
    rho = np.einsum('nij,nkj->nik', X, X.conj())
    rho /= np.trace(rho, axis1=1, axis2=2)[:, None, None]
    return np.real(np.einsum('aij,nij->na', basis_data.conj(), rho))


def tomography_inputs(basis_data, n_particles=400, n_updates=40, seed=99):
    rs = np.random.RandomState(seed)
    d2 = basis_data.shape[0]
    dim = basis_data.shape[1]
    prior = ginibre_coords(rs, n_particles, basis_data)
    true = ginibre_coords(rs, 1, basis_data)[0]
    meas = np.zeros((n_updates, d2))
    ks = rs.randint(1, d2, size=n_updates)
    meas[:, 0] = np.sqrt(dim) / 2
    meas[np.arange(n_updates), ks] = np.sqrt(dim) / 2
    pr1 = np.clip(meas @ true, 0, 1)
    outcomes = (rs.random_sample(n_updates) < pr1).astype(np.int64)
    return dict(prior=prior, meas=meas, outcomes=outcomes, true=true)


# ---------------------------------------------------------------------------
# Drivers
# ---------------------------------------------------------------------------

def _finish(up, log):
    out = dict(
        weights=np.array(up.particle_weights, dtype=float),
        locations=np.array(up.particle_locations, dtype=float),
        normalization_record=np.array([float(np.ravel(v)[0]) for v in up.normalization_record]),
        resample_count=np.int64(up.resample_count),
        min_n_ess=np.float64(up.min_n_ess),
        n_ess=np.float64(up.n_ess),
        est_mean=np.array(up.est_mean(), dtype=float),
        est_cov=np.array(up.est_covariance_mtx(), dtype=float),
        n_events=np.int64(len(log)),
    )
    for i, ev in enumerate(log):
        out['ev%d_w' % i] = ev['w']
        out['ev%d_x' % i] = ev['x']
        out['ev%d_new_x' % i] = ev['new_x']
        for k, v in pack_rng_state(ev['rng']).items():
            out['ev%d_rng_%s' % (i, k)] = v
    return out


def run_precession(ns, inp, seed=0, min_freq=0, a=0.98, trace=None):
    """C1 shape: SimplePrecessionModel, exp-sparse times, default LW."""
    model = ns.SimplePrecessionModel(min_freq=min_freq)
    log = []
    res = recording_resampler(ns, log, a=a)
    np.random.seed(seed)
    up = ns.SMCUpdater(model, inp['prior'].shape[0], FixedPrior(inp['prior']), resampler=res)
    for k in range(len(inp['ts'])):
        up.update(int(inp['outcomes'][k]), np.array([inp['ts'][k]]))
        if trace is not None:
            trace(k, up)
    return _finish(up, log)


def run_rb(ns, inp, seed=0, a=0.98, trace=None):
    """C3 shape: BinomialModel(RandomizedBenchmarkingModel), batch_update(resample_interval=1)."""
    model = ns.BinomialModel(ns.RandomizedBenchmarkingModel())
    log = []
    res = recording_resampler(ns, log, a=a)
    np.random.seed(seed)
    up = ns.SMCUpdater(model, inp['prior'].shape[0], FixedPrior(inp['prior']), resampler=res)
    eps = np.empty((len(inp['ms']),), dtype=model.expparams_dtype)
    eps['m'] = inp['ms']
    eps['n_meas'] = inp['n_meas']
    up.batch_update(inp['counts'], eps, resample_interval=1)
    return _finish(up, log)


def run_tomography(ns, inp, nq=2, seed=0, a=0.98):
    """C4 shape: TomographyModel(pauli_basis(nq)), canonicalize after each resample."""
    model = ns.TomographyModel(ns.pauli_basis(nq))
    log = []
    res = recording_resampler(ns, log, a=a)
    np.random.seed(seed)
    up = ns.SMCUpdater(model, inp['prior'].shape[0], FixedPrior(inp['prior']), resampler=res)
    for k in range(inp['meas'].shape[0]):
        ep = np.empty((1,), dtype=model.expparams_dtype)
        ep['meas'][0] = inp['meas'][k]
        up.update(int(inp['outcomes'][k]), ep)
    return _finish(up, log)


def likelihood_vectors(ns, seed=7):
    """T1 vectors: model.likelihood on random and edge inputs."""
    rs = np.random.RandomState(seed)
    out = {}
    # precession: includes omega=0, large arguments up to ~6e4, outcome label 2 (treated as "not 0")
    x = np.concatenate([rs.random_sample(253), [0.0, 1.0, 0.5]])[:, None]
    ts = np.array([0.0, 1.0, 7.3, (9 / 8.) ** 50, (9 / 8.) ** 99])
    m = ns.SimplePrecessionModel()
    out['prec_x'] = x
    out['prec_t'] = ts
    out['prec_L'] = m.likelihood(np.array([0, 1, 2]), x, ts)
    # RB: edges p in {0,1}, m = 0
    xr = rs.random_sample((256, 3)) * np.array([0.3, 0.5, 0.5]) + np.array([0.7, 0.0, 0.0])
    xr[0] = [1.0, 0.5, 0.5]
    xr[1] = [0.0, 0.5, 0.25]
    xr[2] = [0.99, 0.0, 0.0]
    xr[3] = [0.99, 1.0, 0.0]
    rbm = ns.RandomizedBenchmarkingModel()
    ep = np.empty((4,), dtype=rbm.expparams_dtype)
    ep['m'] = [0, 1, 37, 800]
    out['rb_x'] = xr
    out['rb_m'] = np.array([0, 1, 37, 800], dtype=np.int64)
    out['rb_L'] = rbm.likelihood(np.array([0, 1]), xr, ep)
    out['rb_valid'] = rbm.are_models_valid(np.concatenate([xr, xr * 1.4 - 0.1]))
    out['rb_valid_x'] = np.concatenate([xr, xr * 1.4 - 0.1])
    # Binomial(RB): k in {0, n}, pr1 in {0, 1} via A=0,B=0 / A=1,B=0,p=1
    bm = ns.BinomialModel(ns.RandomizedBenchmarkingModel())
    epb = np.empty((3,), dtype=bm.expparams_dtype)
    epb['m'] = [1, 50, 400]
    epb['n_meas'] = [25, 25, 25]
    ks = np.array([0, 1, 12, 24, 25])
    out['binrb_ks'] = ks
    out['binrb_m'] = np.array([1, 50, 400], dtype=np.int64)
    out['binrb_n'] = np.array([25, 25, 25], dtype=np.int64)
    out['binrb_L'] = bm.likelihood(ks, xr, epb)
    # Binomial(precession): scalar expparam renamed 'x'
    bp = ns.BinomialModel(ns.SimplePrecessionModel())
    epp = np.empty((2,), dtype=bp.expparams_dtype)
    epp['x'] = [3.7, 91.25]
    epp['n_meas'] = [10, 40]
    kp = np.array([0, 3, 10])
    out['binprec_x'] = np.array([3.7, 91.25])
    out['binprec_n'] = np.array([10, 40], dtype=np.int64)
    out['binprec_ks'] = kp
    out['binprec_L'] = bp.likelihood(kp, x, epp)
    # Tomography 1- and 2-qubit: includes clipping on both sides
    for nq in (1, 2):
        basis = ns.pauli_basis(nq)
        tm = ns.TomographyModel(basis)
        d2 = tm.n_modelparams
        xt = ginibre_coords(rs, 128, np.asarray(basis.data))
        xt[:8] *= 3.0                       # unphysical: drives pr1 outside [0, 1]
        ept = np.empty((5,), dtype=tm.expparams_dtype)
        meas = rs.randn(5, d2) * 0.4
        meas[:, 0] = np.sqrt(tm.dim if hasattr(tm, 'dim') else basis.dim) / 2
        ept['meas'] = meas
        out['tomo%d_x' % nq] = xt
        out['tomo%d_meas' % nq] = meas
        out['tomo%d_L' % nq] = tm.likelihood(np.array([0, 1]), xt, ept)
    return out


def canonicalize_vectors(ns, seed=11):
    """T6 vectors: TomographyModel.canonicalize on physical and unphysical coordinates."""
    rs = np.random.RandomState(seed)
    out = {}
    for nq in (1, 2):
        basis = ns.pauli_basis(nq)
        data = np.asarray(basis.data)
        tm = ns.TomographyModel(basis)
        d2 = data.shape[0]
        phys = ginibre_coords(rs, 64, data)
        noisy = phys + 0.15 * rs.randn(64, d2)
        noisy[:, 0] = phys[:, 0] * (1 + 0.05 * rs.randn(64))
        x = np.concatenate([phys, noisy])
        out['canon%d_x' % nq] = x
        out['canon%d_y' % nq] = tm.canonicalize(x.copy())
        tms = ns.TomographyModel(basis, allow_subnormalized=True)
        out['canon%d_y_subnorm' % nq] = tms.canonicalize(x.copy())
    g = ns.gell_mann_basis(3)
    data = np.asarray(g.data)
    tm = ns.TomographyModel(g)
    phys = ginibre_coords(rs, 32, data)
    noisy = phys + 0.2 * rs.randn(32, 9)
    noisy[:, 0] = phys[:, 0]
    x = np.concatenate([phys, noisy])
    out['canon_gm3_x'] = x
    out['canon_gm3_y'] = tm.canonicalize(x.copy())
    return out


def moment_vectors(ns, seed=5):
    """T3 vectors + sqrtm_psd (tests/test_utils.py:132-152 shape)."""
    rs = np.random.RandomState(seed)
    out = {}
    for d in (1, 3, 16):
        n = 777
        x = rs.randn(n, d) * (0.1 + rs.random_sample(d)) + rs.randn(d)
        w = rs.random_sample(n) ** 3
        w[::7] = 0.0
        w /= w.sum()
        pd = ns.ParticleDistribution(particle_locations=x, particle_weights=w)
        out['mom%d_x' % d] = x
        out['mom%d_w' % d] = np.array(pd.particle_weights)
        out['mom%d_mean' % d] = pd.est_mean()
        out['mom%d_cov' % d] = pd.est_covariance_mtx()
        out['mom%d_ness' % d] = np.float64(pd.n_ess)
        S, err = ns.sqrtm_psd(out['mom%d_cov' % d])
        out['mom%d_S' % d] = S
        out['mom%d_Serr' % d] = np.float64(err)
    return out


def design_vectors(ns, seed=21):
    """f2 vectors: SMCUpdater.bayes_risk / expected_information_gain (smc.py:553-657) on clouds with non-uniform
    weights.  The first case is the setting of the reference's known-answer tests (tests/test_metrics.py:40-120):
    BinomialModel(CoinModel()) with Beta(1, 3) particles and n_meas = 1..10."""
    import warnings
    rs = np.random.RandomState(seed)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # (1) coin under BinomialModel, uniform weights, then the same after two data
        n = 3000
        x = rs.beta(1.0, 3.0, size=(n, 1))
        model = ns.BinomialModel(ns.CoinModel())
        up = ns.SMCUpdater(model, n, FixedPrior(x), resample_thresh=0.0)
        ep = np.arange(1, 11, dtype=int).astype(model.expparams_dtype)
        out['coin_x'] = x
        out['coin_nmeas'] = np.arange(1, 11, dtype=np.int64)
        out['coin_risk'] = np.asarray(up.bayes_risk(ep), dtype=float)
        out['coin_ig'] = np.asarray(up.expected_information_gain(ep), dtype=float)
        up.update(2, ep[4:5])
        up.update(1, ep[2:3])
        out['coin_risk_post'] = np.asarray(up.bayes_risk(ep), dtype=float)
        out['coin_ig_post'] = np.asarray(up.expected_information_gain(ep), dtype=float)
        # (2) precession after five updates: several evolution times in one call (n_outcomes constant)
        n = 2500
        x = rs.random_sample((n, 1))
        up = ns.SMCUpdater(ns.SimplePrecessionModel(), n, FixedPrior(x), resample_thresh=0.0)
        for k, (t, o) in enumerate(zip([1.0, 2.3, 4.1, 7.7, 12.9], [0, 1, 0, 0, 1])):
            up.update(o, np.array([t]))
        ts = np.array([0.5, 3.0, 11.0, 40.0, 170.0])
        out['prec_x'] = x
        out['prec_ts'] = ts
        out['prec_risk'] = np.asarray(up.bayes_risk(ts), dtype=float)
        out['prec_ig'] = np.asarray(up.expected_information_gain(ts), dtype=float)
        # (3) randomized benchmarking, d = 3
        n = 2000
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.3 + 0.3 * rs.random_sample(n),
                             0.2 + 0.2 * rs.random_sample(n)])
        rbm = ns.RandomizedBenchmarkingModel()
        up = ns.SMCUpdater(rbm, n, FixedPrior(x), resample_thresh=0.0)
        ep = np.empty((3,), dtype=rbm.expparams_dtype)
        ep['m'] = [1, 20, 150]
        up.update(0, ep[1:2])
        up.update(1, ep[2:3])
        out['rb_x'] = x
        out['rb_m'] = np.array([1, 20, 150], dtype=np.int64)
        out['rb_risk'] = np.asarray(up.bayes_risk(ep), dtype=float)
        out['rb_ig'] = np.asarray(up.expected_information_gain(ep), dtype=float)
        # (4) Binomial(RB): every experiment has its own outcome count
        bm = ns.BinomialModel(ns.RandomizedBenchmarkingModel())
        up = ns.SMCUpdater(bm, n, FixedPrior(x), resample_thresh=0.0)
        epb = np.empty((3,), dtype=bm.expparams_dtype)
        epb['m'] = [5, 60, 300]
        epb['n_meas'] = [4, 12, 25]
        up.update(3, epb[0:1])
        out['binrb_m'] = np.array([5, 60, 300], dtype=np.int64)
        out['binrb_n'] = np.array([4, 12, 25], dtype=np.int64)
        out['binrb_risk'] = np.asarray(up.bayes_risk(epb), dtype=float)
        out['binrb_ig'] = np.asarray(up.expected_information_gain(epb), dtype=float)
        # (5) one-qubit tomography, d = 4 (physical states: no clipping, so the likelihoods are strictly inside (0, 1))
        basis = ns.pauli_basis(1)
        tm = ns.TomographyModel(basis)
        xt = ginibre_coords(rs, 1500, np.asarray(basis.data))
        up = ns.SMCUpdater(tm, 1500, FixedPrior(xt), resample_thresh=0.0)
        ept = np.empty((3,), dtype=tm.expparams_dtype)
        meas = np.zeros((3, 4))
        meas[:, 0] = np.sqrt(2) / 2
        meas[0, 1] = meas[1, 2] = meas[2, 3] = 0.9 * np.sqrt(2) / 2
        ept['meas'] = meas
        up.update(0, ept[0:1])
        up.update(1, ept[2:3])
        out['tomo_x'] = xt
        out['tomo_meas'] = meas
        out['tomo_risk'] = np.asarray(up.bayes_risk(ept), dtype=float)
        out['tomo_ig'] = np.asarray(up.expected_information_gain(ept), dtype=float)
    return out


def mle_vectors(ns, seed=41):
    """f4 vectors: MLEModel (derived_models.py:681-703) — likelihoods raised to a power, alone and over BinomialModel,
    and a short update trajectory without resampling."""
    import warnings
    rs = np.random.RandomState(seed)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = np.concatenate([rs.random_sample(197), [0.0, 1.0, 0.5]])[:, None]
        ts = np.array([0.0, 1.0, 7.3, (9 / 8.) ** 40])
        m = ns.MLEModel(ns.SimplePrecessionModel(), 3.5)
        out['prec_x'], out['prec_t'] = x, ts
        out['prec_L'] = m.likelihood(np.array([0, 1]), x, ts)
        xr = rs.random_sample((200, 3)) * np.array([0.3, 0.5, 0.5]) + np.array([0.7, 0.0, 0.0])
        bm = ns.MLEModel(ns.BinomialModel(ns.RandomizedBenchmarkingModel()), 0.5)
        epb = np.empty((2,), dtype=bm.expparams_dtype)
        epb['m'] = [3, 120]
        epb['n_meas'] = [20, 20]
        out['binrb_x'] = xr
        out['binrb_L'] = bm.likelihood(np.array([0, 7, 20]), xr, epb)
        n = 1500
        prior = rs.random_sample((n, 1))
        up = ns.SMCUpdater(ns.MLEModel(ns.SimplePrecessionModel(), 2.0), n, FixedPrior(prior), resample_thresh=0.0)
        for t, o in zip([0.7, 1.9, 3.3, 6.1, 9.9, 14.2], [0, 1, 0, 0, 1, 0]):
            up.update(o, np.array([t]))
        out['traj_prior'] = prior
        out['traj_w'] = np.array(up.particle_weights)
        out['traj_norm'] = np.array([float(np.ravel(v)[0]) for v in up.normalization_record])
        out['traj_mean'] = np.asarray(up.est_mean(), dtype=float)
    return out


def random_walk_vectors(ns, seed=53):
    """f4 groundwork: the time-dependent decorators (derived_models.py:705-963).  Their update_timestep draws from the
    global NumPy stream, so each trajectory runs under a fixed legacy seed; resampling included."""
    import warnings
    rs = np.random.RandomState(seed)
    out = {}
    n = 1200
    prior = rs.random_sample((n, 1))
    ts = exp_sparse_times(30)
    outcomes = (rs.random_sample(30) >= np.cos(ts * 0.45 / 2) ** 2).astype(int)
    out['prior'], out['ts'], out['outcomes'] = prior, ts, outcomes

    def run(model, prior_arr):
        np.random.seed(7)
        up = ns.SMCUpdater(model, prior_arr.shape[0], FixedPrior(prior_arr))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for t, o in zip(ts, outcomes):
                up.update(int(o), np.array([t]))
        return (np.array(up.particle_locations), np.array(up.particle_weights),
                np.array([float(np.ravel(v)[0]) for v in up.normalization_record]), np.int64(up.resample_count))

    # (1) RandomWalkModel with a fixed normal step distribution
    m = ns.RandomWalkModel(ns.SimplePrecessionModel(), ns.NormalStepDistribution(np.zeros(1), np.array([[1e-6]])))
    out['rw_x'], out['rw_w'], out['rw_norm'], out['rw_rc'] = run(m, prior)
    # (2) GaussianRandomWalkModel, fixed diagonal covariance
    m = ns.GaussianRandomWalkModel(ns.SimplePrecessionModel(), fixed_covariance=np.array([4e-6]))
    out['grw_fixed_x'], out['grw_fixed_w'], out['grw_fixed_norm'], out['grw_fixed_rc'] = run(m, prior)
    # (3) GaussianRandomWalkModel, learned step scale: one extra model parameter sigma >= 0
    prior2 = np.column_stack([prior[:, 0], 2e-3 * rs.random_sample(n)])
    out['prior2'] = prior2
    m = ns.GaussianRandomWalkModel(ns.SimplePrecessionModel())
    out['grw_learn_x'], out['grw_learn_w'], out['grw_learn_norm'], out['grw_learn_rc'] = run(m, prior2)
    out['grw_learn_valid'] = m.are_models_valid(np.array([[0.5, 0.0], [0.5, -1e-9], [-0.1, 1e-3]]))
    # (4) PoisonedModel: ALE (fixed tolerance) and MLE (hedged binomial standard error) noise on the likelihood
    m = ns.PoisonedModel(ns.SimplePrecessionModel(), tol=0.01)
    out['ale_x'], out['ale_w'], out['ale_norm'], out['ale_rc'] = run(m, prior)
    m = ns.PoisonedModel(ns.SimplePrecessionModel(), n_samples=200, hedge=0.5)
    out['mle_x'], out['mle_w'], out['mle_norm'], out['mle_rc'] = run(m, prior)
    return out


def diffusive_vectors(ns, seed=71):
    """f4: DiffusiveTomographyModel (tomography/models.py:228-272), one qubit in the Pauli basis: 5 model parameters
    (4 state coordinates + the diffusion scale), experiments carry a duration ``t``.  Free-running trajectory under a
    fixed legacy seed, resampling and re-canonicalisation included; plus the validity rule and one bare
    update_timestep on fixed particles."""
    import warnings
    rs = np.random.RandomState(seed)
    out = {}
    basis = ns.pauli_basis(1)
    model = ns.DiffusiveTomographyModel(basis)
    bd = np.asarray(basis.data)
    n, n_up = 600, 24
    prior = np.column_stack([ginibre_coords(rs, n, bd), 0.02 + 0.05 * rs.random_sample(n)])
    true = ginibre_coords(rs, 1, bd)[0]
    ep = np.empty((n_up,), dtype=model.expparams_dtype)
    meas = np.zeros((n_up, 4))
    ks = rs.randint(1, 4, size=n_up)
    meas[:, 0] = np.sqrt(2) / 2
    meas[np.arange(n_up), ks] = np.sqrt(2) / 2
    ep['meas'] = meas
    ep['t'] = 0.5 + rs.random_sample(n_up)
    outcomes = (rs.random_sample(n_up) < np.clip(meas @ true, 0, 1)).astype(np.int64)
    out['prior'], out['meas'], out['t'], out['outcomes'] = prior, meas, np.array(ep['t']), outcomes
    out['valid'] = np.asarray(model.are_models_valid(np.array([[0.7, 0.1, 0.0, 0.2, 0.01], [0.7, 0.1, 0.0, 0.2, 0.0],
                                                              [0.7, 0.1, 0.0, 0.2, -0.3]])))
    np.random.seed(11)
    up = ns.SMCUpdater(model, n, FixedPrior(prior))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(n_up):
            up.update(int(outcomes[k]), ep[k:k + 1])
    out['x'] = np.array(up.particle_locations)
    out['w'] = np.array(up.particle_weights)
    out['norm'] = np.array([float(np.ravel(v)[0]) for v in up.normalization_record])
    out['rc'] = np.int64(up.resample_count)
    return out


def mle_design_vectors(ns, seed=29):
    """bayes_risk / expected_information_gain under MLEModel (smc.py:584-591 with derived_models.py:701-703): the last
    outcome's hypothetical likelihood is 1 - sum of the others' POWERED likelihoods."""
    import warnings
    rs = np.random.RandomState(seed)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n = 2000
        x = rs.random_sample((n, 1))
        for tag, gamma in (("g2", 2.0), ("g05", 0.5)):
            up = ns.SMCUpdater(ns.MLEModel(ns.SimplePrecessionModel(), gamma), n, FixedPrior(x), resample_thresh=0.0)
            for t, o in zip([1.0, 2.3, 4.1], [0, 1, 0]):
                up.update(o, np.array([t]))
            ts = np.array([0.5, 3.0, 11.0, 40.0])
            out[tag + '_risk'] = np.asarray(up.bayes_risk(ts), dtype=float)
            out[tag + '_ig'] = np.asarray(up.expected_information_gain(ts), dtype=float)
        out['x'] = x
    return out
