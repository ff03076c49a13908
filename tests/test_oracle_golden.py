"""The oracle must reproduce the UNMODIFIED reference bit for bit.

tests/golden/*.npz were written by tests/golden/make_golden.py from the reference
(QInfer @ 8170c84) run in the build container; here the oracle is driven with
the same seeded inputs and every array is compared exactly.
"""
import warnings

import numpy as np
import pytest

import cases


def _assert_same(got, want, keys=None):
    for k in (keys or want.keys()):
        if k not in got:
            continue
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, k
        assert np.array_equal(a, b, equal_nan=True), "oracle differs from reference on %s" % k


@pytest.fixture(scope="module")
def ns():
    return cases.oracle_namespace()


def _inputs(g, keys):
    return {k: g[k] for k in keys}


def test_precession_c1_bit_exact(ns, golden):
    g = golden("precession_c1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_precession(ns, _inputs(g, ["prior", "ts", "outcomes"]))
    assert int(out["resample_count"]) == int(g["resample_count"]) > 0
    _assert_same(out, g)


def test_precession_minfreq_retry_path_bit_exact(ns, golden):
    g = golden("precession_minfreq")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_precession(ns, _inputs(g, ["prior", "ts", "outcomes"]), min_freq=0.3, a=0.9)
    _assert_same(out, g)


def test_rb_binomial_c3_bit_exact(ns, golden):
    g = golden("rb_binomial_c3")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_rb(ns, _inputs(g, ["prior", "ms", "counts", "n_meas"]))
    _assert_same(out, g)


def test_tomography_c4_bit_exact(ns, golden):
    g = golden("tomography_c4")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_tomography(ns, _inputs(g, ["prior", "meas", "outcomes", "true"]))
    _assert_same(out, g)


def test_likelihood_vectors_bit_exact(ns, golden):
    _assert_same(cases.likelihood_vectors(ns), golden("likelihood_vectors"))


def test_canonicalize_vectors_bit_exact(ns, golden):
    _assert_same(cases.canonicalize_vectors(ns), golden("canonicalize_vectors"))


def test_moment_vectors_bit_exact(ns, golden):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _assert_same(cases.moment_vectors(ns), golden("moment_vectors"))


def test_resample_indices_replay(golden):
    """The stored reference indices are what cumsum + searchsorted give for the stored (w, rng state)."""
    g = golden("precession_c1")
    for i in range(int(g["n_events"])):
        np.random.set_state(cases.unpack_rng_state(g, "ev%d_rng_" % i))
        w = g["ev%d_w" % i]
        u = np.random.random((w.shape[0],))
        js = np.cumsum(w).searchsorted(u, side="right")
        assert np.array_equal(js, g["ev%d_js" % i])
        assert js.max() < w.shape[0]


def test_minfreq_case_exercises_retry_loop(golden):
    """The min_freq fixture must actually hit the postselection retry (prefix-mus quirk, resamplers.py:372)."""
    import smc_oracle as o
    g = golden("precession_minfreq")
    hit = 0
    for i in range(int(g["n_events"])):
        np.random.set_state(cases.unpack_rng_state(g, "ev%d_rng_" % i))
        pd = o.ParticleDistribution(particle_locations=g["ev%d_x" % i], particle_weights=g["ev%d_w" % i])
        pd.particle_weights = g["ev%d_w" % i]      # as the updater holds them (the constructor re-normalises)
        r = o.LiuWestResampler(a=0.9)
        r.record = True
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = r(o.SimplePrecessionModel(min_freq=0.3), pd)
        assert np.array_equal(out.particle_locations, g["ev%d_new_x" % i])
        hit += r.trace["n_iters"] > 1
    assert hit > 0


def test_design_vectors_bit_exact(ns, golden):
    """bayes_risk / expected_information_gain (smc.py:553-657) restated in the oracle, incl. CoinModel."""
    g = golden("design_vectors")
    _assert_same(cases.design_vectors(ns), g)


def test_mle_vectors_bit_exact(ns, golden):
    """MLEModel (derived_models.py:681-703) restated in the oracle."""
    _assert_same(cases.mle_vectors(ns), golden("mle_vectors"))


def test_random_walk_vectors_bit_exact(ns, golden):
    """RandomWalkModel / GaussianRandomWalkModel (derived_models.py:705-963, diagonal covariance) restated in the
    oracle: whole trajectories incl. resampling under a fixed legacy seed.  (Pins the oracle for SURVEY §8 f4; the
    device side of update_timestep is not built yet.)"""
    _assert_same(cases.random_walk_vectors(ns), golden("random_walk_vectors"))


def test_design_known_answers_from_the_reference_tests():
    """tests/test_metrics.py:65-79, 110-120 on the oracle: closed-form Beta-binomial risk (3 decimals) and the
    Mathematica BINOM_IG vector (2 decimals)."""
    import smc_oracle as o
    np.random.seed(0)
    x = np.random.beta(1.0, 3.0, size=(10000, 1))
    model = o.BinomialModel(o.CoinModel())
    up = o.SMCUpdater(model, 10000, cases.FixedPrior(x))
    ep = np.arange(1, 11, dtype=int).astype(model.expparams_dtype)
    exact_risk = 3.0 / (4.0 * 5.0 * (4.0 + ep['n_meas']))
    np.testing.assert_almost_equal(up.bayes_risk(ep), exact_risk, decimal=3)
    binom_ig = np.array([0.104002, 0.189223, 0.261496, 0.324283, 0.379815, 0.429613, 0.474764, 0.516069, 0.554138,
                         0.589446])
    np.testing.assert_almost_equal(up.expected_information_gain(ep), binom_ig, decimal=2)


def test_binomial_pmf_restatement_matches_scipy():
    import smc_oracle as o
    rs = np.random.RandomState(3)
    p = rs.random_sample(500)
    for n, k in [(25, 0), (25, 25), (25, 7), (1, 1), (400, 123)]:
        a = o.binomial_pmf(n, k, p)
        b = o.binomial_pmf_restated(n, k, p)
        np.testing.assert_allclose(b, a, rtol=5e-12, atol=1e-300)


@pytest.mark.refcheck
def test_golden_files_match_a_fresh_reference_run(golden):
    """Build container only: regenerate one case from /root/reference and compare with the committed file."""
    ns_ref = cases.reference_namespace()
    g = golden("precession_c1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_precession(ns_ref, _inputs(g, ["prior", "ts", "outcomes"]))
    _assert_same(out, g)


def test_diffusive_tomography_vectors_bit_exact(ns, golden):
    """f4: DiffusiveTomographyModel (tomography/models.py:228-272) — trajectory with per-update Gaussian diffusion and
    re-canonicalisation, validity rule of the extra parameter."""
    _assert_same(cases.diffusive_vectors(ns), golden("diffusive_vectors"))


def test_mle_design_vectors_bit_exact(ns, golden):
    """Risk and information gain under MLEModel: the last outcome as 1 - sum of the powered others (smc.py:589)."""
    _assert_same(cases.mle_design_vectors(ns), golden("mle_design_vectors"))
