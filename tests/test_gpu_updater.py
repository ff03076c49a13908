"""The reference's own behavioural tests for the hot path, restated against the B200 updater
(SURVEY §4 / T7): tests/test_smc.py:84-137, tests/test_precession_model.py:58-111,
tests/test_distributions.py:615-706, plus the zero-weight policies of smc.py:423-436."""
import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    import qinfer_b200
    return qinfer_b200


@pytest.fixture(scope="module")
def oracle():
    import smc_oracle
    return smc_oracle


def _decimation_setup(qb, n, fraction_zero):
    """The reference drives ESS with a DecimationModel (tests/test_smc.py:55-80) that zeroes the likelihood
    of a chosen fraction of particles.  The same effect with a built-in model: 1-qubit tomography whose
    measurement reads x_1 directly; particles with x_1 = 0 have likelihood exactly 0 for outcome 1."""
    model = qb.TomographyModel(qb.pauli_basis(1))
    x = np.zeros((n, 4))
    x[:, 0] = 1 / np.sqrt(2)
    x[:, 1] = 1.0
    ep = np.empty((1,), dtype=model.expparams_dtype)
    ep['meas'][0] = [0.0, 1.0, 0.0, 0.0]
    return model, x, ep


def test_min_n_ess_is_exactly_4_to_the_k(qb):
    """tests/test_smc.py:122-137: after decimating to 1/4 of the survivors each time, min_n_ess == 4**k."""
    n = 4 ** 6
    model, x, ep = _decimation_setup(qb, n, 0.75)
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resample_thresh=0.0, canonicalize=False)
    alive = n
    for k in reversed(range(1, 6)):
        locs = up.particle_locations.copy()
        live = np.nonzero(locs[:, 1] == 1.0)[0]
        locs[live[live.size // 4:], 1] = 0.0          # keep a quarter of the live particles
        up.particle_locations = locs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            up.update(1, ep)
        alive //= 4
        assert up.n_ess == 4 ** k == alive
        assert up.min_n_ess == 4 ** k


def test_resample_count_and_low_ess_warning(qb):
    """tests/test_smc.py:98-120: ESS below threshold => exactly one resample per update; ESS <= 10 warns."""
    n = 1000
    rs = np.random.RandomState(0)
    x = rs.random_sample((n, 1))
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resample_thresh=1.1)
    np.random.seed(0)
    for k in range(5):
        assert up.resample_count == k
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            up.update(k % 2, np.array([0.7 + k]))
        assert up.just_resampled
    assert up.resample_count == 5
    # low ESS: all the weight on 5 particles
    w = np.zeros(n)
    w[:5] = 0.2
    up.particle_weights = w
    assert up.n_ess == pytest.approx(5.0)
    with pytest.warns(qb.ApproximationWarning, match="Extremely small n_ess"):
        up._maybe_resample()
    with pytest.warns(qb.ResamplerWarning, match="without additional data"):
        up.resample()


def test_precession_estimate_quality(qb):
    """tests/test_precession_model.py:86-110: N=1e4, 100 experiments t in linspace(1,10,100), prior U[0,2],
    true omega = 1: mean to 2 decimals, covariance < 0.01, through batch_update."""
    np.random.seed(0)
    n, true = 10000, 1.0
    model = qb.SimplePrecessionModel()
    prior = qb.UniformDistribution([0, 2])
    up = qb.SMCUpdater(model, n, prior, resampler=qb.LiuWestResampler(), zero_weight_policy='ignore')
    ts = np.linspace(1, 10, 100)
    rs = np.random.RandomState(1)
    outcomes = (rs.random_sample(100) >= np.cos(ts * true / 2) ** 2).astype(int)
    up.batch_update(outcomes, ts, 5)
    assert abs(up.est_mean()[0] - true) < 0.05
    assert up.est_covariance_mtx()[0, 0] < 0.01
    assert len(up.data_record) == 100 and len(up.normalization_record) == 100
    assert np.isfinite(up.log_total_likelihood)
    assert model.call_count == 100 * n


def test_particle_distribution_basics(qb):
    """tests/test_distributions.py:622-706: weights rectified + normalised; n_ess N / 1; moments of an MVN cloud."""
    rs = np.random.RandomState(3)
    n = 100000
    mu = np.array([0.3, -1.0, 2.0])
    A = rs.randn(3, 3)
    cov = A @ A.T / 3
    x = rs.multivariate_normal(mu, cov, size=n)
    pd = qb.ParticleDistribution(particle_locations=x, particle_weights=-np.ones(n) * 3)
    assert np.all(pd.particle_weights > 0) and abs(pd.particle_weights.sum() - 1) < 1e-12
    assert pd.n_ess == pytest.approx(n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(pd.est_mean(), mu, atol=0.02)
        assert np.linalg.norm(pd.est_covariance_mtx() - cov) < 0.05
        corr = pd.est_covariance_mtx(corr=True)
    np.testing.assert_allclose(np.diag(corr), 1.0, atol=1e-12)
    w = np.zeros(n)
    w[17] = 1.0
    pd = qb.ParticleDistribution(particle_locations=x, particle_weights=w)
    assert pd.n_ess == pytest.approx(1.0)
    assert pd.sample(5).shape == (5, 3) and np.all(pd.sample(5) == x[17])


@pytest.mark.parametrize("policy", ["ignore", "skip", "warn", "error", "reset", "bogus"])
def test_zero_weight_policies(qb, policy):
    """smc.py:423-436 with an impossible datum (every particle has likelihood exactly 0)."""
    n = 256
    model, x, ep = _decimation_setup(qb, n, 1.0)
    x[:, 1] = 0.0                                       # pr1 = 0 everywhere: outcome 1 is impossible
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), zero_weight_policy=policy, canonicalize=False)
    before = up.particle_weights.copy()
    if policy == "error":
        with pytest.raises(RuntimeError, match="All particle weights are zero."):
            up.update(1, ep)
        assert np.array_equal(up.particle_weights, before)       # state untouched, as in the reference
        assert up.normalization_record == []
    elif policy == "bogus":
        with pytest.raises(ValueError, match="Invalid zero-weight policy"):
            up.update(1, ep)
    elif policy == "skip":
        up.update(1, ep)
        assert np.array_equal(up.particle_weights, before) and up.normalization_record == []
        assert up.data_record == [1]
    elif policy == "warn":
        with pytest.warns(qb.ApproximationWarning, match="All particle weights are zero"):
            up.update(1, ep, check_for_resample=False)
        assert np.all(up.particle_weights == 0) and up.normalization_record == [0.0]
    elif policy == "reset":
        with pytest.warns(qb.ApproximationWarning, match="Resetting from initial prior"):
            up.update(1, ep, check_for_resample=False)
        assert np.all(up.particle_weights == 0)          # the reference overwrites the reset weights (smc.py:441)
    else:
        up.update(1, ep, check_for_resample=False)
        assert np.all(up.particle_weights == 0)


def test_negative_weights_are_clipped_with_warning(qb):
    """smc.py:416-418.  A tomography 'measurement' with pr1 > 1 is clipped by the model itself, so negative
    weights are injected through the weights attribute instead."""
    n = 128
    model, x, ep = _decimation_setup(qb, n, 0.0)
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), canonicalize=False)
    w = np.full(n, 1.0 / (n - 2))
    w[3] = -1.0 / (n - 2)
    up.particle_weights = w
    with pytest.warns(qb.ApproximationWarning, match="Negative weights"):
        up.update(1, ep, check_for_resample=False)
    got = up.particle_weights
    assert got[3] == 0.0 and np.all(got >= 0) and got.max() <= 1.0


def test_liu_west_zero_covariance_and_standalone_use(qb):
    """resamplers.py:283-294: identical particles => zero-norm covariance => warning + zero_cov_comp."""
    n = 512                                             # powers of two: every partial sum is exact,
    x = np.full((n, 1), 0.375)                          # so E[x^2] - mu^2 is exactly zero
    pd = qb.ParticleDistribution(particle_locations=x, particle_weights=np.ones(n))
    np.random.seed(1)
    with pytest.warns(qb.ResamplerWarning, match="zero norm"):
        out = qb.LiuWestResampler(a=0.9, zero_cov_comp=1e-4)(qb.SimplePrecessionModel(), pd)
    assert isinstance(out, qb.ParticleDistribution) and out.n_particles == n
    spread = out.particle_locations.std()
    assert 0.2 * np.sqrt(1e-4 * (1 - 0.81)) < spread < 5 * np.sqrt(1e-4 * (1 - 0.81))
    # n_particles override
    np.random.seed(1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = qb.LiuWestResampler(a=0.9)(qb.SimplePrecessionModel(), pd, n_particles=123)
    assert out.particle_locations.shape == (123, 1)


def test_liu_west_maxiter_warning_and_no_postselect(qb):
    """resamplers.py:374-381: particles that stay invalid after maxiter iterations produce a warning."""
    n = 400
    rs = np.random.RandomState(0)
    x = 0.1 * rs.random_sample((n, 1))                  # everything below min_freq = 0.5
    pd = qb.ParticleDistribution(particle_locations=x, particle_weights=np.ones(n))
    np.random.seed(2)
    with pytest.warns(qb.ResamplerWarning, match="failed to find valid models"):
        qb.LiuWestResampler(maxiter=3)(qb.SimplePrecessionModel(min_freq=0.5), pd)
    np.random.seed(2)
    r = qb.LiuWestResampler(postselect=False)
    with warnings.catch_warnings():
        warnings.simplefilter("error", qb.ResamplerWarning)
        r(qb.SimplePrecessionModel(min_freq=0.5), pd)
    assert r.last_n_iters == 1


def test_hypothetical_update_shapes_and_values(qb):
    """smc.py:324-386: (n_outcomes, n_expparams, n_particles) weights, normalisations, likelihoods."""
    import smc_oracle as o
    n = 300
    rs = np.random.RandomState(4)
    x = rs.random_sample((n, 1))
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x))
    ou = o.SMCUpdater(o.SimplePrecessionModel(), n, cases.FixedPrior(x))
    ts = np.array([0.5, 2.0, 9.0])
    w, L, norm = up.hypothetical_update(np.array([0, 1]), ts, return_likelihood=True, return_normalization=True)
    w0, L0, norm0 = ou.hypothetical_update(np.array([0, 1]), ts, return_likelihood=True, return_normalization=True)
    assert w.shape == (2, 3, n) and L.shape == (2, 3, n) and norm.shape == (2, 3, 1)
    # The reference forms L(1) = 1 - fl(cos^2) (abstract_model.py:679): one ulp of pr0 ~ 1 is an ABSOLUTE error
    # of ~2e-16 in L, i.e. up to 1e-10 relative when L ~ 1e-6 (the kernel's sin^2 is the accurate one — checked
    # against long double).  So: likelihoods to 4e-16 absolute, weights to that floor scaled by prior / norm.
    np.testing.assert_allclose(L, L0, rtol=1e-12, atol=4e-16)
    floor = 4e-16 * (1.0 / n) / norm0
    assert np.all(np.abs(w - w0) <= 1e-12 * np.abs(w0) + floor)
    np.testing.assert_allclose(norm, norm0, rtol=1e-12)


def test_reset_and_setters(qb):
    n = 64
    np.random.seed(5)
    up = qb.SMCUpdater(qb.RandomizedBenchmarkingModel(), n, qb.UniformDistribution([[0.8, 1], [0.3, 0.5], [0.3, 0.5]]))
    assert up.particle_locations.shape == (n, 3) and up.n_rvs == 3 and up.n_particles == n
    assert up.n_ess == pytest.approx(n)
    with pytest.raises(ValueError):
        up.reset(n_particles=10, only_params=[0])
    old = up.particle_locations.copy()
    up.reset(only_params=[0])
    new = up.particle_locations
    assert np.array_equal(new[:, 1:], old[:, 1:]) and not np.array_equal(new[:, 0], old[:, 0])
    with pytest.raises(qb.UnsupportedModelError):
        qb.SMCUpdater(object(), n, qb.UniformDistribution([0, 1]))
    with pytest.raises(ValueError, match="Both a resample_a"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            qb.SMCUpdater(qb.SimplePrecessionModel(), n, qb.UniformDistribution([0, 1]), resample_a=0.9,
                          resampler=qb.LiuWestResampler())


def _run_modes(qb, model_factory, n, x, steps, modes, seed=3):
    """Run the same update sequence under several (lazy, fuse) modes; return the end states."""
    out = []
    for lazy, fuse in modes:
        np.random.seed(seed)
        up = qb.SMCUpdater(model_factory(), n, cases.FixedPrior(x), lazy=lazy, fuse=fuse)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for outcome, ep, chk in steps:
                up.update(outcome, ep, check_for_resample=chk)
        out.append(dict(rc=up.resample_count, rec=np.array(up.normalization_record), w=up.particle_weights.copy(),
                        x=up.particle_locations.copy(), min_ess=up.min_n_ess, ess=up.n_ess,
                        launches=up._cloud.update_launches))
    return out


def test_lazy_pipelining_is_bit_identical_to_eager(qb):
    """lazy=True, fuse=1 launches update k+1 speculatively behind update k (device-side guard); the trajectory,
    records, resample count and final cloud must equal the call-by-call (reference-semantics) run exactly."""
    n = 50000
    rs = np.random.RandomState(8)
    x = rs.random_sample((n, 1))
    ts = (9.0 / 8.0) ** np.arange(70)
    outcomes = (rs.random_sample(70) >= np.cos(ts * 0.5 / 2) ** 2).astype(int)
    steps = [(int(outcomes[k]), ts[k:k + 1], (k % 7 != 3)) for k in range(70)]
    a, b = _run_modes(qb, qb.SimplePrecessionModel, n, x, steps, [(False, None), (True, 1)])
    assert a["rc"] == b["rc"] and a["rc"] >= 3
    assert np.array_equal(a["rec"], b["rec"]) and np.array_equal(a["w"], b["w"]) and np.array_equal(a["x"], b["x"])
    assert a["min_ess"] == b["min_ess"] and a["ess"] == b["ess"]


@pytest.mark.parametrize("fuse", [2, 5, 8])
def test_fused_updates_match_one_by_one(qb, fuse):
    """K consecutive updates in one launch: same resample decisions at the same steps, records
    and weights equal to the one-by-one path up to the omitted intermediate renormalisation (a few ulp);
    a mid-batch resample trigger rolls the batch back to the triggering step."""
    n = 40000
    rs = np.random.RandomState(21)
    x = rs.random_sample((n, 1))
    ts = (9.0 / 8.0) ** np.arange(60)
    outcomes = (rs.random_sample(60) >= np.cos(ts * 0.5 / 2) ** 2).astype(int)
    steps = [(int(outcomes[k]), ts[k:k + 1], True) for k in range(60)]
    # resampling draws from the legacy NumPy stream: identical decisions => identical streams => same clouds
    a, b = _run_modes(qb, qb.SimplePrecessionModel, n, x, steps, [(False, None), (True, fuse)])
    assert a["rc"] == b["rc"] and a["rc"] >= 3
    assert b["launches"] < a["launches"]
    np.testing.assert_allclose(b["rec"], a["rec"], rtol=1e-12)
    np.testing.assert_allclose(b["x"], a["x"], rtol=1e-9, atol=1e-12)
    # after several resamples the clouds differ by ~1e-13 relative (moments computed from weights a few ulp
    # apart); cos^2(t x / 2) with t ~ 1e3 turns that into ~1e-9 on individual weights
    np.testing.assert_allclose(b["w"], a["w"], rtol=1e-6, atol=1e-12 * a["w"].max())
    assert b["min_ess"] == pytest.approx(a["min_ess"], rel=1e-11) and b["ess"] == pytest.approx(a["ess"], rel=1e-11)


def test_fused_batch_update_rb(qb):
    """batch_update(resample_interval) through fused launches vs the oracle's loop (RB o Binomial, d = 3)."""
    import smc_oracle as o
    n = 3000
    inp = cases.rb_inputs(n_particles=n, n_updates=40, seed=77)
    res = []
    for ns in (qb, o):
        model = ns.BinomialModel(ns.RandomizedBenchmarkingModel())
        np.random.seed(1)
        up = ns.SMCUpdater(model, n, cases.FixedPrior(inp['prior']))
        eps = np.empty((40,), dtype=model.expparams_dtype)
        eps['m'] = inp['ms']
        eps['n_meas'] = inp['n_meas']
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            up.batch_update(inp['counts'], eps, resample_interval=5)
        res.append((up.resample_count, np.array([float(np.ravel(v)[0]) for v in up.normalization_record]),
                    up.est_mean(), up.est_covariance_mtx(), up.min_n_ess))
    (rc_g, rec_g, m_g, c_g, e_g), (rc_o, rec_o, m_o, c_o, e_o) = res
    assert rc_g == rc_o
    np.testing.assert_allclose(rec_g, rec_o, rtol=1e-9)
    np.testing.assert_allclose(m_g, m_o, rtol=1e-6)
    np.testing.assert_allclose(c_g, c_o, rtol=1e-5, atol=1e-12)
    assert e_g == pytest.approx(e_o, rel=1e-8)


@pytest.mark.parametrize("policy", ["skip", "error", "warn"])
def test_lazy_mode_zero_weight_events(qb, policy):
    """A zero-weight event under lazy=True: the speculative successor cancels itself, state stays coherent."""
    n = 256
    model, x, ep = _decimation_setup(qb, n, 1.0)
    x[: n // 2, 1] = 0.0                                # half the particles make outcome 1 impossible
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), zero_weight_policy=policy, canonicalize=False, lazy=True,
                       resample_thresh=0.0)
    ep0 = np.empty((1,), dtype=model.expparams_dtype)
    ep0['meas'][0] = [0.0, 1.0, 0.0, 0.0]
    up.update(1, ep0, check_for_resample=False)          # kills half the particles
    locs = up.particle_locations.copy()
    locs[:, 1] = 0.0
    up.particle_locations = locs                         # now outcome 1 is impossible for every survivor
    w_before = up.particle_weights.copy()
    up.update(1, ep0, check_for_resample=False)          # zero-weight event (pending)
    if policy == "error":
        with pytest.raises(RuntimeError, match="All particle weights are zero."):
            up.update(0, ep0, check_for_resample=False)  # settles the pending step first
        assert np.array_equal(up.particle_weights, w_before)
    elif policy == "skip":
        up.update(0, ep0, check_for_resample=False)      # the skipped step leaves the weights; this one applies
        w = up.particle_weights
        assert len(up.normalization_record) == 2 and abs(w.sum() - 1) < 1e-12
        assert np.array_equal(w > 0, w_before > 0)
    else:
        with pytest.warns(qb.ApproximationWarning, match="All particle weights are zero"):
            up.update(0, ep0, check_for_resample=False)
            w = up.particle_weights
        assert np.all(w == 0)


def test_simple_est_prec_and_rb_match_the_oracle(qb, tmp_path):
    """qinfer.simple_est's one-call estimators (simple_est.py:141-260) on the GPU engine vs the same recipe driven
    through the oracle: identical legacy seed => identical prior draw, resample decisions and estimates."""
    import smc_oracle as o
    rs = np.random.RandomState(17)
    # precession: rows of (counts, t, n_shots), also through a CSV file
    ts = np.linspace(0.5, 40, 60)
    n_shots = 30
    counts = rs.binomial(n_shots, 1 - np.cos(ts * 0.62 / 2) ** 2)
    table = np.column_stack([counts, ts, np.full_like(ts, n_shots)])
    csv = tmp_path / "prec.csv"
    np.savetxt(str(csv), table, delimiter=",", fmt=["%d", "%.17g", "%d"])
    np.random.seed(4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, var, extra = qb.simple_est_prec(table, freq_max=1.0, n_particles=5000, return_all=True)
    assert extra['updater'].resample_count > 0
    np.random.seed(4)
    model = o.BinomialModel(o.SimplePrecessionModel(0.0))
    ou = o.SMCUpdater(model, 5000, o.UniformDistribution([0, 1.0]))
    eps = np.empty((60,), dtype=model.expparams_dtype)
    eps['x'], eps['n_meas'] = ts, n_shots
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ou.batch_update(counts, eps, resample_interval=1)
    assert extra['updater'].resample_count == ou.resample_count
    assert mean == pytest.approx(ou.est_mean()[0], rel=1e-6)
    assert var == pytest.approx(ou.est_covariance_mtx()[0, 0], rel=1e-5)
    assert abs(mean - 0.62) < 0.02
    np.random.seed(4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean2, var2 = qb.simple_est_prec(str(csv), freq_max=1.0, n_particles=5000)
    assert mean2 == pytest.approx(mean, rel=1e-12)
    # randomized benchmarking: rows of (counts, m, n_shots)
    ms = np.linspace(1, 400, 50).astype(int)
    counts = rs.binomial(25, 0.5 * 0.99 ** ms + 0.5)
    table = np.column_stack([counts, ms, np.full_like(ms, 25)])
    np.random.seed(9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, cov = qb.simple_est_rb(table, p_min=0.8, n_particles=6000)
    np.random.seed(9)
    model = o.BinomialModel(o.RandomizedBenchmarkingModel())
    prior = o.PostselectedDistribution(o.UniformDistribution([[0.8, 1.0], [0, 1], [0, 1]]), model)
    ou = o.SMCUpdater(model, 6000, prior)
    eps = np.empty((50,), dtype=model.expparams_dtype)
    eps['m'], eps['n_meas'] = ms, 25
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ou.batch_update(counts, eps, resample_interval=1)
    np.testing.assert_allclose(mean, ou.est_mean(), rtol=1e-6)
    np.testing.assert_allclose(cov, ou.est_covariance_mtx(), rtol=1e-4, atol=1e-10)
    assert abs(mean[0] - 0.99) < 0.01


def test_read_side_estimators_match_the_oracle(qb, oracle):
    """est_entropy / est_credible_region / est_meanfn / sample (distributions.py:320-333, 411-430, 457-465, 557-596)
    on a weighted cloud, against the oracle's ParticleDistribution holding the same particles."""
    rs = np.random.RandomState(8)
    n = 3000
    x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    m_g, m_o = qb.RandomizedBenchmarkingModel(), oracle.RandomizedBenchmarkingModel()
    up = qb.SMCUpdater(m_g, n, cases.FixedPrior(x), resample_thresh=0.0)
    ou = oracle.SMCUpdater(m_o, n, cases.FixedPrior(x), resample_thresh=0.0)
    ep = np.empty((2,), dtype=m_g.expparams_dtype)
    ep['m'] = [7, 60]
    for u in (up, ou):
        u.update(0, ep[0:1])
        u.update(1, ep[1:2])
    w_o, x_o = ou.particle_weights, ou.particle_locations
    np.testing.assert_allclose(up.particle_weights, w_o, rtol=1e-12)
    nz = w_o[w_o > 0]
    assert up.est_entropy() == pytest.approx(-np.sum(np.log(nz) * nz), rel=1e-12)
    inside, outside = up.est_credible_region(level=0.9, return_outside=True, modelparam_slice=slice(0, 2))
    order = np.argsort(w_o)[::-1]
    k = int(np.sum(np.cumsum(w_o[order]) <= 0.9)) + 1
    assert inside.shape == (k, 2) and outside.shape == (n - k, 2)
    assert np.array_equal(np.sort(inside[:, 0]), np.sort(x_o[order][:k, 0]))
    assert up.est_meanfn(lambda l: l[:, 0] ** 2) == pytest.approx(ou.est_meanfn(lambda l: l[:, 0] ** 2), rel=1e-12)
    faces, verts = up.region_est_hull(level=0.5, modelparam_slice=slice(0, 2))
    assert faces.shape[1:] == (2, 2) and verts.shape[1] == 2 and verts.shape[0] >= 3
    np.random.seed(4)
    s_g = up.sample(n=500)
    np.random.seed(4)
    s_o = ou.sample(n=500)
    assert s_g.shape == (500, 3)
    assert np.mean(np.all(s_g == s_o, axis=1)) > 0.99          # same uniforms, same CDF up to the weights' last bits


@pytest.mark.parametrize("n,level", [(3000, 0.9), (10 ** 6, 0.95), (10 ** 6, 0.5), (200001, 0.999)])
def test_f3_credible_region_on_the_device(qb, oracle, n, level):
    """est_credible_region (distributions.py:558-614) by radix selection + compaction on the device: the same member
    set as the reference's argsort + cumsum, sorted by weight, and only the members are downloaded (D2H bytes
    proportional to the region, not to the cloud)."""
    rs = np.random.RandomState(12)
    x = rs.random_sample((n, 1))
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resample_thresh=0.0)
    for t, o in zip([1.1, 2.7, 6.3, 11.9, 23.0], [0, 1, 0, 0, 1]):
        up.update(o, np.array([t]))
    w = up.particle_weights
    order = np.argsort(w)[::-1]
    k = int(np.sum(np.cumsum(w[order]) <= level)) + 1
    region = up.est_credible_region(level=level)
    assert region.shape == (k, 1)
    assert up._cloud.last_d2h_bytes == 8 * k * 2 and k < n
    want = x[order][:k]
    wk = w[order][k - 1]
    assert np.sum(w == wk) == 1                      # (distinct weights at the boundary: the member set is unique)
    assert np.array_equal(np.sort(region[:, 0]), np.sort(want[:, 0]))
    # sorted by weight, heaviest first: compare where the weights are distinct
    distinct = np.concatenate([[True], np.diff(w[order][:k]) != 0]) & np.concatenate([np.diff(w[order][:k]) != 0, [True]])
    assert np.array_equal(region[distinct], want[distinct])
    # slicing and the uniform-weights tie case (every weight equal: any K of them, lowest index first)
    up2 = qb.SMCUpdater(qb.RandomizedBenchmarkingModel(), 5000, cases.FixedPrior(rs.random_sample((5000, 3))))
    reg = up2.est_credible_region(level=0.25, modelparam_slice=slice(1, 3))
    assert reg.shape == (int(np.sum(np.cumsum(np.full(5000, 1 / 5000)) <= 0.25)) + 1, 2)
    with pytest.raises(IndexError):
        up2.est_credible_region(level=1.5)


def test_f3_meanfn_and_entropy_on_the_device(qb, oracle):
    """est_meanfn (distributions.py:411-430): a function written with operators is evaluated on the device tensor and
    reduced there (moment kernel); a NumPy-only function falls back to the reference's host expression; est_entropy
    (distributions.py:457-465) is a device reduction."""
    import torch
    rs = np.random.RandomState(3)
    n = 400000
    x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    up = qb.SMCUpdater(qb.RandomizedBenchmarkingModel(), n, cases.FixedPrior(x), resample_thresh=0.0)
    ep = np.empty((2,), dtype=up.model.expparams_dtype)
    ep['m'] = [7, 60]
    up.update(0, ep[0:1])
    up.update(1, ep[1:2])
    w = up.particle_weights
    calls = []

    def poly(l):
        calls.append(type(l))
        return l ** 2 + 3.0 * l

    got = up.est_meanfn(poly)
    assert calls == [torch.Tensor]                                   # evaluated on the device tensor only
    np.testing.assert_allclose(got, np.einsum('i,ij->j', w, x ** 2 + 3.0 * x), rtol=1e-12)
    got1 = up.est_meanfn(lambda l: l[:, 0] * l[:, 2])
    assert np.shape(got1) == () and got1 == pytest.approx(float(np.dot(w, x[:, 0] * x[:, 2])), rel=1e-12)
    host = up.est_meanfn(lambda l: np.sin(np.asarray(l)[:, 1]))      # needs NumPy: the reference's host route
    assert host == pytest.approx(float(np.dot(w, np.sin(x[:, 1]))), rel=1e-12)
    nz = w[w > 0]
    assert up.est_entropy() == pytest.approx(-np.sum(np.log(nz) * nz), rel=1e-11)
