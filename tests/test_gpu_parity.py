"""GPU parity tests: every CUDA kernel of the hot path, called through the C ABI /
plugin surface, against (a) the golden vectors produced by the unmodified
reference and (b) the NumPy oracle on seeded inputs.

Bars (BASELINE.json north_star): resample indices bit-exact; posterior
mean/covariance within 1e-6 relative; the per-kernel tolerances below are much
tighter and are written next to each assertion.  Measured maxima are appended to
gpurun_out/parity_report.txt so they can be quoted.
"""
import os
import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report(name, value):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write("%s %r\n" % (name, value))


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


@pytest.fixture(scope="module")
def qb():
    import qinfer_b200
    return qinfer_b200


@pytest.fixture(scope="module")
def oracle():
    import smc_oracle
    return smc_oracle


def gpu_namespace(qb, **lw_kwargs):
    return cases.Namespace(
        name="b200", SMCUpdater=qb.SMCUpdater, LiuWestResampler=qb.LiuWestResampler,
        ParticleDistribution=qb.ParticleDistribution, SimplePrecessionModel=qb.SimplePrecessionModel,
        SimpleInversionModel=qb.SimpleInversionModel, RandomizedBenchmarkingModel=qb.RandomizedBenchmarkingModel,
        BinomialModel=qb.BinomialModel, CoinModel=qb.CoinModel, MLEModel=qb.MLEModel,
        TomographyModel=qb.TomographyModel, DiffusiveTomographyModel=qb.DiffusiveTomographyModel,
        RandomWalkModel=qb.RandomWalkModel, GaussianRandomWalkModel=qb.GaussianRandomWalkModel,
        PoisonedModel=qb.PoisonedModel, NormalStepDistribution=qb.MultivariateNormalDistribution,
        pauli_basis=qb.pauli_basis,
        gell_mann_basis=qb.gell_mann_basis, UniformDistribution=qb.UniformDistribution,
        PostselectedDistribution=qb.PostselectedDistribution, sqrtm_psd=qb.sqrtm_psd)


# ---------------------------------------------------------------------------
# T1 — likelihood kernels vs the reference's Model.likelihood (golden vectors)
# ---------------------------------------------------------------------------
def test_t1_likelihood_vectors(qb, golden):
    g = golden("likelihood_vectors")
    got = cases.likelihood_vectors(gpu_namespace(qb))
    # cos/pow/log/exp differ from glibc in the last ulp or two; 1 - pr0 near pr0 ~ 1 turns that into
    # an absolute error of a few 1e-16, hence the absolute floor (and k*log(pr1) amplifies it by k/pr1
    # under BinomialModel(SimplePrecessionModel), hence its 1e-14).
    for key, rtol, atol in [("prec_L", 1e-12, 2e-15), ("rb_L", 1e-12, 2e-15), ("binrb_L", 2e-12, 1e-300),
                            ("binprec_L", 2e-12, 1e-14), ("tomo1_L", 1e-13, 1e-15), ("tomo2_L", 1e-13, 1e-15)]:
        assert got[key].shape == g[key].shape
        np.testing.assert_allclose(got[key], g[key], rtol=rtol, atol=atol, err_msg=key)
        report("t1_" + key + "_max_abs", float(np.max(np.abs(got[key] - g[key]))))
    assert np.array_equal(got["rb_valid"], g["rb_valid"])          # validity masks are exact


def test_t1_binomial_extremes(qb, oracle):
    """k in {0, n}, pr1 exactly 0 and exactly 1 (A=B=0; A=1,B=0,p=1), large n_meas."""
    x = np.array([[1.0, 0.0, 0.0], [1.0, 1.0, 0.0], [0.9, 0.5, 0.25], [0.99, 0.3, 0.6]])
    m_b, m_o = qb.BinomialModel(qb.RandomizedBenchmarkingModel()), oracle.BinomialModel(
        oracle.RandomizedBenchmarkingModel())
    ep = np.empty((2,), dtype=m_b.expparams_dtype)
    ep['m'] = [1, 10]
    ep['n_meas'] = [30, 2000]
    ks = np.array([0, 1, 30])
    got = m_b.likelihood(ks, x, ep)
    want = m_o.likelihood(ks, x, ep)
    np.testing.assert_allclose(got, want, rtol=5e-11, atol=1e-300)
    assert m_b.call_count == ks.size * x.shape[0] * ep.size        # abstract_model.py:466-468


# ---------------------------------------------------------------------------
# T2 — fused update vs SMCUpdater.hypothetical_update / update
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 1000, 1024, 100003, 3 * 10 ** 6 + 5])
def test_t2_fused_update_precession(qb, oracle, n):
    rs = np.random.RandomState(n % 1000)
    x = rs.random_sample((n, 1))
    w = rs.random_sample(n) ** 2
    w /= w.sum()
    t = np.array([17.3])
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), zero_weight_policy='ignore')
    up.particle_weights = w
    for outcome in (0, 1):
        L = oracle.SimplePrecessionModel().likelihood(np.array([outcome]), x, t)[0, :, 0]
        hyp = w * L
        want_norm = np.sum(hyp)
        want_w = hyp / want_norm
        up.update(outcome, t, check_for_resample=False)
        got_w = up.particle_weights
        # rtol 1e-12 wherever the weight matters; particles sitting on a zero of cos^2 (L ~ 1e-13) see the
        # libm-vs-CUDA difference of the reduced argument relatively larger, hence the absolute floor
        np.testing.assert_allclose(got_w, want_w, rtol=1e-12, atol=1e-15 * want_w.max())
        assert abs(up.normalization_record[-1] - want_norm) <= 1e-12 * want_norm
        assert abs(up.n_ess - 1 / np.sum(want_w ** 2)) <= 1e-10 * up.n_ess
        report("t2_prec_n%d_o%d_w_rel" % (n, outcome), relerr(got_w, want_w, 1e-300))
        w = want_w
        up.particle_weights = w


def test_t2_fused_update_all_models_trajectory(qb, oracle, golden):
    """Several consecutive lazy-normalised updates (no resampling) for each model family."""
    rs = np.random.RandomState(5)
    runs = []
    # RB o Binomial (d = 3)
    g = golden("rb_binomial_c3")
    mb, mo = qb.BinomialModel(qb.RandomizedBenchmarkingModel()), oracle.BinomialModel(
        oracle.RandomizedBenchmarkingModel())
    eps = np.empty((6,), dtype=mb.expparams_dtype)
    eps['m'] = g["ms"][:6]
    eps['n_meas'] = int(g["n_meas"])
    runs.append((mb, mo, g["prior"], [(int(g["counts"][k]), eps[k:k + 1]) for k in range(6)]))
    # tomography 2 qubits (d = 16)
    g = golden("tomography_c4")
    tb, to = qb.TomographyModel(qb.pauli_basis(2)), oracle.TomographyModel(oracle.pauli_basis(2))
    steps = []
    for k in range(8):
        ep = np.empty((1,), dtype=tb.expparams_dtype)
        ep['meas'][0] = g["meas"][k]
        steps.append((int(g["outcomes"][k]), ep))
    runs.append((tb, to, g["prior"], steps))
    # interleaved RB (d = 4)
    xi = np.column_stack([0.9 + 0.1 * rs.random_sample(500), 0.9 + 0.1 * rs.random_sample(500),
                          0.5 * rs.random_sample(500), 0.5 * rs.random_sample(500)])
    ib, io = qb.RandomizedBenchmarkingModel(interleaved=True), oracle.RandomizedBenchmarkingModel(interleaved=True)
    steps = []
    for k in range(6):
        ep = np.empty((1,), dtype=ib.expparams_dtype)
        ep['m'] = 5 + 20 * k
        ep['reference'] = bool(k % 2)
        steps.append((k % 2, ep))
    runs.append((ib, io, xi, steps))
    # Binomial(precession): scalar expparam renamed 'x'
    pb, po = qb.BinomialModel(qb.SimplePrecessionModel()), oracle.BinomialModel(oracle.SimplePrecessionModel())
    steps = []
    for k in range(6):
        ep = np.empty((1,), dtype=pb.expparams_dtype)
        ep['x'] = 1.5 ** k
        ep['n_meas'] = 20
        steps.append((3 + 2 * k, ep))
    runs.append((pb, po, rs.random_sample((777, 1)), steps))

    for mb, mo, prior, steps in runs:
        n = prior.shape[0]
        gb = qb.SMCUpdater(mb, n, cases.FixedPrior(prior), canonicalize=False)
        ob = oracle.SMCUpdater(mo, n, cases.FixedPrior(prior), canonicalize=False)
        for outcome, ep in steps:
            gb.update(outcome, ep, check_for_resample=False)
            ob.update(outcome, ep, check_for_resample=False)
            # absolute floor: under BinomialModel, k*log(pr1) turns the last-ulp difference between CUDA's and
            # glibc's cos into ~1e-16 * k / pr1 relative, which only shows on particles of negligible weight
            ow = ob.particle_weights
            np.testing.assert_allclose(gb.particle_weights, ow, rtol=2e-11, atol=1e-15 * ow.max())
            np.testing.assert_allclose(gb.normalization_record[-1], np.ravel(ob.normalization_record[-1])[0],
                                       rtol=1e-11)
            np.testing.assert_allclose(gb.n_ess, ob.n_ess, rtol=1e-10)
        report("t2_traj_%s_w_rel" % type(mb).__name__, relerr(gb.particle_weights, ob.particle_weights, 1e-300))
        assert gb.min_n_ess == pytest.approx(ob.min_n_ess, rel=1e-10)
        assert mb.call_count == len(steps) * n


# ---------------------------------------------------------------------------
# T3 — moments vs ParticleDistribution.est_mean / est_covariance_mtx (golden)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 3, 16])
def test_t3_moments_golden(qb, golden, d):
    g = golden("moment_vectors")
    pd = qb.ParticleDistribution(particle_locations=g["mom%d_x" % d], particle_weights=g["mom%d_w" % d])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, cov = pd.est_mean(), pd.est_covariance_mtx()
    np.testing.assert_allclose(mean, g["mom%d_mean" % d], rtol=1e-12, atol=1e-14)
    scale = np.max(np.abs(g["mom%d_cov" % d]))
    np.testing.assert_allclose(cov, g["mom%d_cov" % d], rtol=0, atol=1e-12 * max(scale, 1.0))   # north_star: 1e-6
    assert pd.n_ess == pytest.approx(float(g["mom%d_ness" % d]), rel=1e-12)
    report("t3_d%d_cov_abs" % d, float(np.max(np.abs(cov - g["mom%d_cov" % d]))))


@pytest.mark.parametrize("d,n", [(1, 10 ** 6 + 3), (2, 33), (3, 400001), (4, 1000), (5, 7777), (9, 5000),
                                 (16, 250007), (16, 3), (64, 2000)])
def test_t3_moments_oracle_sizes(qb, oracle, d, n):
    """Every moments kernel (register, DMMA d=16, generic) incl. ragged tails, vs np.dot / einsum."""
    from qinfer_b200.engine import host_moments
    rs = np.random.RandomState(d * 7 + 1)
    x = rs.randn(n, d) * 0.3 + np.arange(d) * 0.1
    w = rs.random_sample(n)
    w /= w.sum()
    sw, mean, m2 = host_moments(w, x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_mean, want_cov = oracle.particle_mean(w, x), oracle.particle_covariance_mtx(w, x)
    cov = m2 - np.outer(mean, mean)
    assert sw == pytest.approx(1.0, abs=1e-12)
    np.testing.assert_allclose(mean, want_mean, rtol=1e-11, atol=1e-13)
    # uncentred formula: error ~ eps * |mu|^2, far inside the 1e-6 bar here
    np.testing.assert_allclose(cov, want_cov, rtol=0, atol=1e-11 * (1 + np.max(np.abs(want_mean)) ** 2))
    assert np.array_equal(m2, m2.T)


# ---------------------------------------------------------------------------
# T4 — CDF + draw: bit-identical to np.cumsum(w).searchsorted(u, 'right') (clamped)
# ---------------------------------------------------------------------------
def _draw_on_gpu(qb, w, u, scan):
    from qinfer_b200 import _lib
    from qinfer_b200.engine import DeviceCloud
    import torch
    n = w.shape[0]
    cloud = DeviceCloud(qb.describe_model(qb.SimplePrecessionModel()), n)
    cloud.upload_locations(np.zeros((n, 1)))
    cloud.upload_weights(w)
    cloud._resample_scratch(u.shape[0])
    cdf = cloud.cdf(_lib.QB_SCAN_EXACT if scan == 'exact' else _lib.QB_SCAN_FAST)
    if scan == 'exact' and n >= 32768:
        _draw_on_gpu.fell_back = cloud.exact_scan_fell_back()
    cloud._u.copy_(torch.from_numpy(u))
    js = cloud.draw(cloud._u, u.shape[0])
    n_bad, overflow = cloud.read_counter()
    return cdf.cpu().numpy(), js.cpu().numpy(), overflow


def _weight_cases():
    rs = np.random.RandomState(42)
    out = {}
    w = rs.random_sample(5000); out["random"] = w / w.sum()
    w = rs.random_sample(5000); w[rs.random_sample(5000) < 0.6] = 0.0; out["zeros"] = w / w.sum()
    out["ties"] = np.full(4096, 1.0 / 4096)                       # exact ties / exactly representable steps
    w = np.full(3000, 1e-310); w[1234] = 1.0; out["denormal_plus_onehot"] = w
    w = np.zeros(777); w[5] = 1.0; out["onehot"] = w
    w = rs.random_sample(100003) ** 8; out["skewed_100k"] = w / w.sum()
    w = np.exp(-0.5 * ((np.arange(10 ** 6) - 3e5) / 2e4) ** 2) + 1e-30; out["gaussian_1m"] = w / w.sum()
    w = rs.random_sample(1); out["single"] = w / w.sum()
    return out


@pytest.mark.parametrize("name", sorted(_weight_cases().keys()))
def test_t4_exact_scan_and_draw_bit_identical(qb, name):
    w = _weight_cases()[name]
    rs = np.random.RandomState(len(name))
    u = rs.random_sample(max(w.shape[0], 2000))
    cdf, js, overflow = _draw_on_gpu(qb, w, u, 'exact')
    want_cdf = np.cumsum(w)
    assert np.array_equal(cdf, want_cdf), "sequential-order scan differs from np.cumsum"
    want = want_cdf.searchsorted(u, side='right')
    assert overflow == int(np.sum(want >= w.shape[0]))
    assert np.array_equal(js, np.minimum(want, w.shape[0] - 1))


def _big_weight_cases():
    rs = np.random.RandomState(7)
    out = {}
    w = rs.random_sample(32768 + 5); out["random_33k"] = w / w.sum()
    w = rs.random_sample(10 ** 6 + 3) ** 6; out["skewed_1m"] = w / w.sum()
    w = rs.random_sample(3 * 10 ** 6); w[rs.random_sample(w.size) < 0.7] = 0.0; out["zeros_3m"] = w / w.sum()
    out["ties_1m"] = np.full(2 ** 20, 2.0 ** -20)                    # every partial sum exact
    w = np.full(2 ** 20 + 17, 1.0 / 3.0); out["ties_third"] = w / w.sum()   # identical weights: many exact ties
    w = np.zeros(500000); w[400000:] = rs.random_sample(100000); out["leading_zeros"] = w / w.sum()
    w = np.full(300000, 1e-312); w[123456] = 1.0; w[250000] = 0.5; out["denormals_and_jumps"] = w
    w = np.exp(-0.5 * ((np.arange(10 ** 7) - 6e6) / 3e5) ** 2) + 1e-40; out["gaussian_10m"] = w / w.sum()
    w = rs.random_sample(10 ** 7); out["random_10m"] = w / w.sum()
    w = 2.0 ** -rs.randint(1, 60, size=200000).astype(float); out["powers_of_two"] = w / 1.0
    return out


@pytest.mark.parametrize("name", sorted(_big_weight_cases().keys()))
def test_t4_parallel_exact_scan_bit_identical(qb, name):
    """The binade-integer replay scan (n >= 32768) equals np.cumsum bit for bit, and so do the drawn indices."""
    w = _big_weight_cases()[name]
    rs = np.random.RandomState(len(name) + 1)
    u = rs.random_sample(min(w.shape[0], 10 ** 6)) * min(1.0, float(np.sum(w)))
    cdf, js, overflow = _draw_on_gpu(qb, w, u, 'exact')
    want_cdf = np.cumsum(w)
    nbad = int(np.sum(cdf != want_cdf))
    assert nbad == 0, "%d of %d CDF entries differ from np.cumsum (first at %d)" % (
        nbad, w.size, int(np.argmax(cdf != want_cdf)))
    assert _draw_on_gpu.fell_back == 0, "the parallel replay scan handed over to the sequential kernel"
    want = want_cdf.searchsorted(u, side='right')
    assert np.array_equal(js, np.minimum(want, w.shape[0] - 1))


@pytest.mark.parametrize("name", ["skewed_1m", "zeros_3m", "leading_zeros", "denormals_and_jumps", "powers_of_two",
                                  "ties_third", "random_33k"])
def test_t4_chained_exact_scan_equals_cumsum_of_the_concatenation(qb, name):
    """qb_cdf_chained (SURVEY §8e parity mode): slabs scanned one after the other, each continuing the running fp64
    sum from the last CDF entry of the slab before it, concatenate to np.cumsum of the whole weight vector bit for
    bit — through the parallel replay scan (slabs >= 32768) and the one-lane kernel (small slabs) alike."""
    import torch
    from qinfer_b200 import _lib
    from qinfer_b200.engine import DeviceCloud
    w = _big_weight_cases()[name]
    n = w.shape[0]
    cuts = [0, int(0.37 * n) + 1, int(0.37 * n) + 1 + 1000, int(0.8 * n) + 3, n]      # uneven, one small slab
    desc = qb.describe_model(qb.SimplePrecessionModel())
    carry = torch.zeros((1,), dtype=torch.float64, device="cuda")
    pieces = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        cloud = DeviceCloud(desc, hi - lo)
        cloud.upload_locations(np.zeros((hi - lo, 1)))
        cloud.upload_weights(w[lo:hi])
        cloud.stats[_lib.QB_STAT_INV_NORM] = 1.0             # slab weights are normalised GLOBALLY already
        cdf = cloud.cdf(_lib.QB_SCAN_EXACT, carry=carry)
        if hi - lo >= 32768:
            assert cloud.exact_scan_fell_back() == 0
        pieces.append(cdf.cpu().numpy())
        carry = cdf[-1:].clone()
    got, want = np.concatenate(pieces), np.cumsum(w)
    nbad = int(np.sum(got != want))
    assert nbad == 0, "%d of %d chained CDF entries differ from np.cumsum (first at %d)" % (
        nbad, n, int(np.argmax(got != want)))


def test_t4_exact_scan_falls_back_on_negative_weights(qb):
    """Weights outside the replay's model (negative) are detected and the sequential kernel takes over."""
    rs = np.random.RandomState(2)
    w = rs.random_sample(70000) / 70000
    w[12345] = -1e-6
    cdf, js, _ = _draw_on_gpu(qb, w, rs.random_sample(1000) * 0.9, 'exact')
    assert _draw_on_gpu.fell_back == 1
    assert np.array_equal(cdf, np.cumsum(w))


def test_t4_exact_scan_with_lazy_normalisation(qb):
    """The scan runs on w[i] * inv_norm (the deferred normalisation), like the updater uses it."""
    from qinfer_b200 import _lib
    n = 400000
    rs = np.random.RandomState(3)
    x = rs.random_sample((n, 1))
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resample_thresh=0.0)
    for k in range(5):
        up.update(k % 2, np.array([1.7 ** k]))
    w = up.particle_weights                       # normalised weights as the host sees them (w * inv_norm)
    cloud = up._cloud
    cloud._resample_scratch(n)
    cdf = cloud.cdf(_lib.QB_SCAN_EXACT).cpu().numpy()
    assert np.array_equal(cdf, np.cumsum(w))


def test_t4_draw_edges(qb):
    """u exactly on a CDF step goes right (side='right'); u above cdf[-1] is clamped and counted."""
    w = np.array([0.25, 0.25, 0.0, 0.25, 0.125])                  # cdf[-1] = 0.875 < 1
    u = np.array([0.0, 0.25, 0.5, 0.4999999999999999, 0.75, 0.874, 0.875, 0.99])
    cdf, js, overflow = _draw_on_gpu(qb, w, u, 'exact')
    want = np.cumsum(w).searchsorted(u, side='right')
    assert list(want) == [0, 1, 3, 1, 4, 4, 5, 5]
    assert overflow == 2
    assert np.array_equal(js, np.minimum(want, 4))


def test_t4_fast_scan_is_close_and_indices_rarely_differ(qb):
    rs = np.random.RandomState(9)
    w = rs.random_sample(10 ** 6)
    w /= w.sum()
    u = rs.random_sample(10 ** 6)
    cdf, js, _ = _draw_on_gpu(qb, w, u, 'fast')
    want_cdf = np.cumsum(w)
    assert np.max(np.abs(cdf - want_cdf)) < 1e-12
    want = np.minimum(want_cdf.searchsorted(u, side='right'), w.shape[0] - 1)
    mism = int(np.sum(js != want))
    report("t4_fast_scan_index_mismatches_of_1e6", mism)
    assert mism <= 5 and np.all(np.abs(js - want) <= 1)


# ---------------------------------------------------------------------------
# T5 — Liu-West: teacher-forced replay of the reference's recorded resample events
# ---------------------------------------------------------------------------
def _replay_events(qb, g, model, a):
    """Teacher-forced: same weights, locations, legacy-RNG state and (mean, cov) as the reference saw."""
    import smc_oracle as o
    worst = 0.0
    for i in range(int(g["n_events"])):
        w, x, new_x = g["ev%d_w" % i], g["ev%d_x" % i], g["ev%d_new_x" % i]
        pd = qb.ParticleDistribution(particle_locations=x, particle_weights=w)
        pd.particle_weights = w                    # exactly the weights the reference's updater held
        res = qb.LiuWestResampler(a=a)
        np.random.set_state(cases.unpack_rng_state(g, "ev%d_rng_" % i))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # moments as NumPy computes them (the device moments are checked in T3); sqrtm_psd of a
            # rank-deficient covariance turns 1e-17 noise into 1e-9, so the move is compared on equal inputs
            mean, cov = o.particle_mean(w, x), o.particle_covariance_mtx(w, x)
            out = res(model, pd, precomputed_mean=mean, precomputed_cov=cov)
        assert res.last_overflow == 0
        got = out.particle_locations
        assert got.shape == new_x.shape
        worst = max(worst, relerr(got, new_x, 1e-12))
        # d = 1 is bit-exact by construction; d > 1 differs only by BLAS's summation order in S @ eps
        np.testing.assert_allclose(got, new_x, rtol=1e-12, atol=1e-14, err_msg="event %d" % i)
    return worst


def test_t5_resample_indices_bit_exact_all_cases(qb, golden):
    """Same weights + same legacy-RNG state => the device draw returns the reference's indices bit for bit."""
    for case in ("precession_c1", "precession_minfreq", "rb_binomial_c3", "tomography_c4"):
        g = golden(case)
        for i in range(int(g["n_events"])):
            w, u = g["ev%d_w" % i], g["ev%d_u" % i]
            cdf, js, overflow = _draw_on_gpu(qb, w, u, 'exact')
            assert overflow == 0
            assert np.array_equal(js, g["ev%d_js" % i]), "%s event %d" % (case, i)


def test_t5_liu_west_events_precession(qb, golden):
    report("t5_prec_locs_rel", _replay_events(qb, golden("precession_c1"), qb.SimplePrecessionModel(), 0.98))


def test_t5_liu_west_retry_quirk(qb, golden):
    """min_freq > 0 forces the postselection retry loop; parity needs the prefix-`mus` quirk (resamplers.py:372)."""
    g = golden("precession_minfreq")
    report("t5_minfreq_locs_rel", _replay_events(qb, g, qb.SimplePrecessionModel(min_freq=0.3), 0.9))


def test_t5_liu_west_events_rb(qb, golden):
    report("t5_rb_locs_rel", _replay_events(qb, golden("rb_binomial_c3"),
                                            qb.BinomialModel(qb.RandomizedBenchmarkingModel()), 0.98))


def test_t5_liu_west_events_tomography(qb, golden):
    # the recorded new_x are pre-canonicalize (the resampler's own output)
    report("t5_tomo_locs_rel", _replay_events(qb, golden("tomography_c4"), qb.TomographyModel(qb.pauli_basis(2)),
                                              0.98))


@pytest.mark.parametrize("kind,n", [("precession", 1000), ("precession", 4096), ("rb", 3000)])
def test_t5_small_cloud_single_launch_resample_equals_the_oracle(qb, kind, n):
    """Parity mode on a small cloud (qb_lw_small_resample: ONE single-CTA launch on host-drawn np.random variates):
    resample indices bit-identical to the oracle's (np.cumsum + searchsorted), locations 1e-12, the same number of
    retry iterations, and the legacy stream left at the same position; and identical to the staged launches."""
    import smc_oracle as oracle
    rs = np.random.RandomState(n)
    if kind == "precession":
        gm, om = qb.SimplePrecessionModel(min_freq=0.35), oracle.SimplePrecessionModel(min_freq=0.35)
        x = 0.3 + 0.4 * rs.random_sample((n, 1))
    else:
        gm, om = qb.RandomizedBenchmarkingModel(), oracle.RandomizedBenchmarkingModel()
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.6 * rs.random_sample(n), 0.4 * rs.random_sample(n)])
    w = rs.random_sample(n) ** 4
    w[rs.randint(0, n, n // 7)] = 0.0
    w /= w.sum()
    np.random.seed(31)
    ores = oracle.LiuWestResampler(a=0.9)
    ores.record = True
    want = ores(om, oracle.ParticleDistribution(particle_locations=x, particle_weights=w)).particle_locations
    want_next = np.random.random()
    results = {}
    for small in (True, False):
        np.random.seed(31)
        res = qb.LiuWestResampler(a=0.9, rng='numpy', scan='exact')
        res._small = small
        up = qb.SMCUpdater(gm, n, cases.FixedPrior(x), resampler=res)
        up.particle_weights = w
        launches0 = up._cloud.launches
        up.resample()
        results[small] = (up._cloud._js.cpu().numpy().copy(), up.particle_locations.copy(), res.last_n_iters,
                          np.random.random(), up._cloud.launches - launches0, up.particle_weights.copy())
    js, locs, iters, nxt, launches, wts = results[True]
    assert np.array_equal(js, ores.trace["js"])
    assert iters == ores.trace["n_iters"] and iters > 2                     # the retry loop ran, equally long
    assert nxt == want_next
    np.testing.assert_allclose(locs, want, rtol=1e-12, atol=1e-14)
    assert np.all(np.asarray(gm.are_models_valid(locs)))
    assert np.all(wts == 1.0 / n)
    js2, locs2, iters2, nxt2, launches2, _ = results[False]
    assert np.array_equal(js, js2) and iters == iters2 and nxt == nxt2
    np.testing.assert_allclose(locs, locs2, rtol=1e-12, atol=1e-14)
    assert launches < launches2                                             # fewer launches than the staged path
    report("t5_small_%s_%d_locs_rel" % (kind, n), relerr(locs, want))


def test_t5_small_cloud_zero_covariance_warns_like_the_reference(qb):
    import warnings as _w
    n = 512
    x = np.full((n, 1), 0.5)
    np.random.seed(3)
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x),
                       resampler=qb.LiuWestResampler(rng='numpy', scan='exact'))
    with _w.catch_warnings(record=True) as rec:
        _w.simplefilter("always")
        up.resample()
    assert any("zero norm" in str(r.message) for r in rec)
    np.random.seed(3)
    np.random.random((n,))
    eps = np.random.randn(1, n)
    h = np.sqrt(1 - 0.98 ** 2)
    want = (0.98 * 0.5 + (1 - 0.98) * 0.5) + (h * np.sqrt(1e-10)) * eps[0]
    np.testing.assert_allclose(up.particle_locations[:, 0], want, rtol=1e-13)


# ---------------------------------------------------------------------------
# T6 — tomography canonicalize vs TomographyModel.canonicalize (golden)
# ---------------------------------------------------------------------------
def test_t6_canonicalize_golden(qb, golden):
    g = golden("canonicalize_vectors")
    for nq in (1, 2):
        x = g["canon%d_x" % nq]
        got = qb.TomographyModel(qb.pauli_basis(nq)).canonicalize(x.copy())
        np.testing.assert_allclose(got, g["canon%d_y" % nq], rtol=0, atol=1e-10)
        report("t6_pauli%d_abs" % nq, float(np.max(np.abs(got - g["canon%d_y" % nq]))))
        got = qb.TomographyModel(qb.pauli_basis(nq), allow_subnormalized=True).canonicalize(x.copy())
        np.testing.assert_allclose(got, g["canon%d_y_subnorm" % nq], rtol=0, atol=1e-10)
        # physical (already PSD) inputs are passed through untouched when subnormalised states are allowed
        assert np.array_equal(got[:64], x[:64])
    got = qb.TomographyModel(qb.gell_mann_basis(3)).canonicalize(g["canon_gm3_x"].copy())
    np.testing.assert_allclose(got, g["canon_gm3_y"], rtol=0, atol=1e-10)


def test_t6_canonicalize_output_is_physical(qb):
    """tests/base_test.py:413-420 (test_canonicalize): canonical states are valid density operators."""
    rs = np.random.RandomState(2)
    basis = qb.pauli_basis(2)
    x = cases.ginibre_coords(rs, 5000, basis.data) + 0.2 * rs.randn(5000, 16)
    x[:, 0] = 0.5
    y = qb.TomographyModel(basis).canonicalize(x)
    rho = np.einsum('na,aij->nij', y, basis.data)
    ev = np.linalg.eigvalsh(rho)
    assert ev.min() > -1e-12
    np.testing.assert_allclose(np.trace(rho, axis1=1, axis2=2).real, 1.0, atol=1e-12)


# ---------------------------------------------------------------------------
# T8 — free-running trajectories vs the reference (golden), legacy RNG seed 0
# ---------------------------------------------------------------------------
def _check_trajectory(name, out, g, mean_rtol=1e-6):
    assert int(out["resample_count"]) == int(g["resample_count"])
    np.testing.assert_allclose(out["normalization_record"], g["normalization_record"], rtol=1e-7)
    np.testing.assert_allclose(out["est_mean"], g["est_mean"], rtol=mean_rtol)          # north_star: 1e-6
    # cov = E[xx^T] - mu mu^T cancels catastrophically once |mu|^2/|cov| ~ 1e10 (the reference's own
    # formula is then only good to eps*kappa ~ 1e-5 relative), so the 1e-6 bar gets an absolute floor of a few
    # ulps of |mu|^2
    floor = 16 * np.spacing(1.0) * float(np.max(np.abs(g["est_mean"])) ** 2)
    np.testing.assert_allclose(out["est_cov"], g["est_cov"], rtol=1e-6, atol=floor)
    np.testing.assert_allclose(out["min_n_ess"], g["min_n_ess"], rtol=1e-7)
    report("t8_%s_mean_rel" % name, relerr(out["est_mean"], g["est_mean"], 1e-300))
    report("t8_%s_locs_rel" % name, relerr(out["locations"], g["locations"], 1e-12))


def test_t8_precession_c1_free_running(qb, golden):
    g = golden("precession_c1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_precession(gpu_namespace(qb), {k: g[k] for k in ("prior", "ts", "outcomes")})
    _check_trajectory("prec_c1", out, g)
    # the recorded resample events saw the same indices as the reference did
    for i in range(int(g["n_events"])):
        np.testing.assert_allclose(out["ev%d_new_x" % i], g["ev%d_new_x" % i], rtol=1e-6, atol=1e-9)


def test_t8_precession_minfreq_free_running(qb, golden):
    g = golden("precession_minfreq")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_precession(gpu_namespace(qb), {k: g[k] for k in ("prior", "ts", "outcomes")},
                                   min_freq=0.3, a=0.9)
    _check_trajectory("prec_minfreq", out, g)


def test_t8_rb_binomial_c3_free_running(qb, golden):
    g = golden("rb_binomial_c3")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_rb(gpu_namespace(qb), {k: g[k] for k in ("prior", "ms", "counts", "n_meas")})
    _check_trajectory("rb_c3", out, g)


def test_t8_tomography_c4_free_running(qb, golden):
    g = golden("tomography_c4")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cases.run_tomography(gpu_namespace(qb), {k: g[k] for k in ("prior", "meas", "outcomes", "true")})
    _check_trajectory("tomo_c4", out, g)


# ---------------------------------------------------------------------------
# Device RNG (throughput mode): known-answer for Philox4x32-10 + moments of the output
# ---------------------------------------------------------------------------
def test_philox_uniform_matches_numpy_philox(qb):
    """Cross-check against NumPy's independent Philox4x64? No: 4x32 is not in NumPy — use the Random123
    known-answer vector for Philox4x32-10 (counter 0, key 0) and a host restatement for a stream."""
    import torch
    from qinfer_b200.engine import DeviceCloud
    cloud = DeviceCloud(qb.describe_model(qb.SimplePrecessionModel()), 16)
    out = torch.empty((8,), dtype=torch.float64, device=cloud.device)
    cloud.rng_uniform(out, 8, 0, 0)
    got = out.cpu().numpy()
    want = _philox_uniform_host(8, 0, 0)
    assert np.array_equal(got, want)
    # Random123 KAT: philox4x32_10(ctr=0, key=0) = 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    r = _philox4x32_10(np.zeros(4, dtype=np.uint64), np.zeros(2, dtype=np.uint64))
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    big = torch.empty((10 ** 6,), dtype=torch.float64, device=cloud.device)
    cloud.rng_uniform(big, 10 ** 6, 1234, 77)
    b = big.cpu().numpy()
    assert 0.0 <= b.min() and b.max() < 1.0 and abs(b.mean() - 0.5) < 2e-3 and abs(b.var() - 1 / 12) < 1e-3
    cloud.rng_normal(big, 10 ** 6, 1234, 77)
    z = big.cpu().numpy()
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3 and abs((z ** 4).mean() - 3) < 5e-2


def _philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(v) for v in ctr]
    k = [int(v) for v in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def _philox_uniform_host(n, seed, offset):
    out = np.empty(n)
    for p in range((n + 1) // 2):
        ctr = offset + p
        r = _philox4x32_10([ctr & 0xFFFFFFFF, ctr >> 32, 0, 0], [seed & 0xFFFFFFFF, seed >> 32])
        vals = [((r[0] >> 5) * 67108864.0 + (r[1] >> 6)) / 9007199254740992.0,
                ((r[2] >> 5) * 67108864.0 + (r[3] >> 6)) / 9007199254740992.0]
        out[2 * p] = vals[0]
        if 2 * p + 1 < n:
            out[2 * p + 1] = vals[1]
    return out


def test_philox_mode_resampling_preserves_moments(qb):
    """Throughput mode (device RNG + fast scan): Liu-West preserves mean and covariance in expectation."""
    n = 400000
    rs = np.random.RandomState(1)
    x = rs.random_sample((n, 1))
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=7, scan='fast')
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resampler=res)
    for k in range(12):
        up.update(k % 2, np.array([1.3 ** k]), check_for_resample=False)
    m0, c0 = up.est_mean(), up.est_covariance_mtx()
    up.resample()
    m1, c1 = up.est_mean(), up.est_covariance_mtx()
    assert up.n_ess == pytest.approx(n)
    sigma = np.sqrt(c0[0, 0])
    assert abs(m1[0] - m0[0]) < 6 * sigma / np.sqrt(up.min_n_ess)
    assert abs(c1[0, 0] / c0[0, 0] - 1) < 0.05


# ---------------------------------------------------------------------------
# Device MT19937 (parity mode at scale): NumPy's legacy global stream continued on the GPU
# ---------------------------------------------------------------------------
def _mt_cloud(qb):
    from qinfer_b200.engine import DeviceCloud
    return DeviceCloud(qb.describe_model(qb.SimplePrecessionModel()), 16)


def _same_state(a, b):
    return (a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
            and (a[3] == 0 or a[4] == b[4]))


@pytest.mark.parametrize("seed,skip,n", [(0, 0, 1), (1, 0, 5), (2, 3, 311), (3, 0, 312), (4, 1, 313), (5, 623, 10007),
                                         (6, 17, 2 * 10 ** 6 + 1), (7, 0, 10 ** 7)])
def test_mt19937_uniform_bit_identical_to_numpy(qb, seed, skip, n):
    """np.random.random((n,)) (resamplers.py:319) regenerated on the device from np.random's own state: same doubles,
    same generator state afterwards, and the host stream continues seamlessly."""
    import torch
    cloud = _mt_cloud(qb)
    np.random.seed(seed)
    if skip:
        np.random.randint(0, 2 ** 31, size=skip)      # move the position inside the 624-word block
    s0 = np.random.get_state()
    want = np.random.random((n,))
    s1 = np.random.get_state()
    tail_want = np.random.random((7,))
    np.random.set_state(s0)
    out = torch.empty((n,), dtype=torch.float64, device=cloud.device)
    cloud.mt19937_uniform(out, n)
    got = out.cpu().numpy()
    assert np.array_equal(got, want)
    assert _same_state(np.random.get_state(), s1)
    assert np.array_equal(np.random.random((7,)), tail_want)


@pytest.mark.parametrize("seed,pre,m", [(0, 0, 1), (1, 0, 2), (2, 1, 3), (3, 0, 10), (4, 1, 1000), (5, 0, 1001),
                                        (6, 1, 1), (7, 1, 2), (8, 0, 10 ** 6 + 1), (9, 1, 4 * 10 ** 6)])
def test_mt19937_normal_matches_numpy_stream(qb, seed, pre, m):
    """np.random.randn (resamplers.py:332; legacy polar method with a cached variate): the accepted candidates, the
    words consumed and the final state (incl. has_gauss / cached) are exact; values agree to 1 ulp of log()."""
    import torch
    cloud = _mt_cloud(qb)
    np.random.seed(seed)
    if pre:
        np.random.randn(pre)                           # leaves a cached second variate behind when pre is odd
    s0 = np.random.get_state()
    want = np.random.randn(m)
    s1 = np.random.get_state()
    tail_want = np.random.randn(5)
    np.random.set_state(s0)
    out = torch.empty((m,), dtype=torch.float64, device=cloud.device)
    cloud.mt19937_normal(out, m)
    got = out.cpu().numpy()
    s1_got = np.random.get_state()
    assert s1_got[0] == s1[0] and np.array_equal(s1_got[1], s1[1]) and s1_got[2:4] == s1[2:4]
    if s1[3]:
        assert abs(s1_got[4] - s1[4]) <= 4 * np.spacing(abs(s1[4]))
    ulps = np.abs(got - want) / np.spacing(np.abs(want))
    report("mt19937_normal_max_ulp_m%d" % m, float(ulps.max()))
    report("mt19937_normal_frac_identical_m%d" % m, float(np.mean(got == want)))
    assert ulps.max() <= 4
    np.testing.assert_allclose(np.random.randn(5), tail_want, rtol=1e-15)


def test_mt19937_mode_reproduces_golden_trajectories(qb, golden):
    """rng='mt19937' draws the same stream as rng='numpy' (the reference's), so the free-running golden
    trajectories — resample count, records, estimates — are reproduced with the variates made on the device."""
    class MtLiuWest(qb.LiuWestResampler):
        def __init__(self, *a, **k):
            k.setdefault('rng', 'mt19937')
            super(MtLiuWest, self).__init__(*a, **k)

    ns = gpu_namespace(qb)
    ns.LiuWestResampler = MtLiuWest
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = golden("precession_c1")
        _check_trajectory("mt_prec_c1", cases.run_precession(ns, {k: g[k] for k in ("prior", "ts", "outcomes")}), g)
        g = golden("rb_binomial_c3")
        _check_trajectory("mt_rb_c3", cases.run_rb(ns, {k: g[k] for k in ("prior", "ms", "counts", "n_meas")}), g)
        g = golden("tomography_c4")
        _check_trajectory("mt_tomo_c4",
                          cases.run_tomography(ns, {k: g[k] for k in ("prior", "meas", "outcomes", "true")}), g)


def test_mt19937_mode_same_indices_as_numpy_mode_at_scale(qb):
    """At 2e6 particles the device-generated stream picks exactly the particles the host stream picks."""
    n = 2 * 10 ** 6
    x = np.random.RandomState(3).random_sample((n, 1))
    outs = []
    for mode in ("numpy", "mt19937"):
        up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x),
                           resampler=qb.LiuWestResampler(a=0.98, rng=mode))
        for k in range(10):
            up.update(k % 2, np.array([1.4 ** k]), check_for_resample=False)
        np.random.seed(1234)
        up.resample()
        outs.append((up.particle_locations.copy(), np.random.get_state()))
    (xa, sa), (xb, sb) = outs
    assert np.array_equal(sa[1], sb[1]) and sa[2:4] == sb[2:4]
    np.testing.assert_allclose(xb, xa, rtol=0, atol=1e-15)


# ---------------------------------------------------------------------------
# Fused draw + move (device-RNG mode): one kernel == the four staged launches, bit for bit
# ---------------------------------------------------------------------------
def _fused_case(qb, kind, n, seed):
    rs = np.random.RandomState(seed)
    if kind == "prec":
        model, x = qb.SimplePrecessionModel(), rs.random_sample((n, 1))
    elif kind == "prec_minfreq":          # a tenth of the offspring is invalid -> several retry rounds
        model, x = qb.SimplePrecessionModel(min_freq=0.35), 0.35 + 0.3 * rs.random_sample((n, 1))
    elif kind == "rb":                    # d = 3, many offspring violate A + B <= 1
        model = qb.RandomizedBenchmarkingModel()
        x = np.column_stack([0.8 + 0.2 * rs.random_sample(n), 0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    else:                                 # interleaved RB, d = 4
        model = qb.RandomizedBenchmarkingModel(interleaved=True)
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.8 + 0.2 * rs.random_sample(n),
                             0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    w = rs.random_sample(n) ** 4
    w[rs.randint(0, n, size=max(n // 50, 1))] = 0.0          # exact zeros: repeated CDF entries
    if n > 10:
        w[n // 3] = 40.0 * w.sum() / n                       # one heavy particle: a long bucket range in the guide
    return model, x, w / w.sum()


@pytest.mark.parametrize("kind,n,n_new", [("prec", 1000, None), ("prec", 4096, None), ("prec", 100003, None),
                                          ("prec_minfreq", 50001, None), ("rb", 65537, None), ("rb", 30000, 29999),
                                          ("rb_il", 20001, None), ("prec", 3000001, None), ("prec", 9, None),
                                          ("rb", 2 ** 20, None), ("prec", 10 ** 7, None)])
def test_fused_draw_move_bit_identical_to_staged(qb, kind, n, n_new):
    """qb_cdf(FAST_GUIDE) + qb_lw_draw_move / qb_lw_draw_retry against qb_cdf(FAST) + qb_rng_uniform + qb_draw +
    qb_rng_normal + qb_lw_move / qb_compact_invalid / qb_lw_retry on the same cloud and Philox streams: identical
    particles, identical retry rounds, identical clamp count."""
    model, x, w = _fused_case(qb, kind, n, 17)
    out = []
    for fused in (False, True):
        res = qb.LiuWestResampler(a=0.9, rng='philox', seed=1234567, scan='fast', draw='guided')
        res._fused = fused
        up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
        up.particle_weights = w
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            new = res(model, up, n_particles=n_new)
        out.append((new.particle_locations.copy(), res.last_n_iters, res.last_overflow, res._philox_offset))
    (xs, it_s, ov_s, off_s), (xf, it_f, ov_f, off_f) = out
    assert xs.shape == xf.shape == ((n if n_new is None else n_new), x.shape[1])
    assert (it_s, ov_s, off_s) == (it_f, ov_f, off_f)
    assert np.array_equal(xs, xf)
    if kind in ("prec_minfreq", "rb", "rb_il"):
        assert it_f > 1                                        # the retry kernel was exercised


@pytest.mark.parametrize("n", [4096, 10 ** 5 + 3, 2 ** 21])
def test_fused_resample_against_oracle_given_the_same_variates(qb, oracle, n):
    """The fused kernel regenerates element i of the Philox streams; materialise the same streams with
    qb_rng_uniform / qb_rng_normal, hand them to the NumPy oracle's Liu-West arithmetic, compare the particles."""
    import torch
    from qinfer_b200.engine import _ptr, _stream
    model, x, w = _fused_case(qb, "prec", n, 5)
    seed = 99
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=seed, scan='fast', draw='guided')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    u = torch.empty((n,), dtype=torch.float64, device=cloud.device)
    e = torch.empty((n,), dtype=torch.float64, device=cloud.device)
    cloud.rng_uniform(u, n, seed, 0)
    cloud.rng_normal(e, n, seed ^ 0x9E3779B97F4A7C15, (n + 1) // 2)
    cdf = cloud.cdf(qb._lib.QB_SCAN_FAST).cpu().numpy().copy()
    mean, cov = up.est_mean(), up.est_covariance_mtx()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        new = res(model, up)
    got = new.particle_locations
    js = np.minimum(cdf.searchsorted(u.cpu().numpy(), side='right'), n - 1)     # resamplers.py:318-321
    S = np.real(res.h * oracle.sqrtm_psd(cov)[0])
    mus = 0.98 * x[js] + (1 - 0.98) * mean                                      # resamplers.py:325
    want = mus + np.dot(S, e.cpu().numpy()[None, :]).T                          # resamplers.py:332
    valid = want[:, 0] > 0
    assert valid.mean() > 0.95
    assert np.array_equal(got[valid], want[valid])                              # d = 1: bit-exact by construction
    # the re-associated scan ends at ~1 and is non-decreasing up to the rounding of a partial-sum boundary
    # (a zero weight next to a thread boundary may step back by an ulp; the guide scatter tolerates that)
    assert np.all(np.diff(cdf) >= -4.5e-16) and abs(cdf[-1] - 1) < 1e-12



# ---------------------------------------------------------------------------
# f2 — bayes_risk / expected_information_gain (smc.py:553-657) reduced on the device
# ---------------------------------------------------------------------------
def test_design_vectors_against_the_reference(qb, golden):
    """Golden vectors of the UNMODIFIED reference (coin/binomial, precession, RB, binomial RB, tomography).  The
    device reduces each hypothetical posterior to (N, first and second moment about the mean); tolerance 1e-10
    relative: summation order and the shifted-moment formula differ from the reference's two-pass einsum."""
    g = golden("design_vectors")
    got = cases.design_vectors(gpu_namespace(qb))
    for key in sorted(g):
        if key.endswith(("risk", "ig", "risk_post", "ig_post")):
            assert got[key].shape == g[key].shape, key
            # the reference forms the last outcome's likelihood as 1 - sum(others): under BinomialModel that cancels
            # to +-1e-16 and log() of it gives NaN (binrb_ig[2]); the device evaluates the pmf and stays finite
            assert np.all(np.isfinite(got[key])), key
            ok = ~np.isnan(g[key])
            np.testing.assert_allclose(got[key][ok], g[key][ok], rtol=1e-10, atol=0, err_msg=key)
            report("f2_" + key + "_rel", relerr(got[key][ok], g[key][ok]))


def test_design_known_answers_of_the_reference_tests(qb):
    """tests/test_metrics.py:65-79 (risk, 3 decimals against the closed form) and :110-120 (BINOM_IG, 2 decimals),
    restated with the reference's sizes: 10 000 Beta(1, 3) particles, BinomialModel(CoinModel()), n_meas = 1..10."""
    np.random.seed(0)
    x = np.random.beta(1.0, 3.0, size=(10000, 1))
    model = qb.BinomialModel(qb.CoinModel())
    up = qb.SMCUpdater(model, 10000, cases.FixedPrior(x))
    ep = np.arange(1, 11, dtype=int).astype(model.expparams_dtype)
    a, b = 1.0, 3.0
    exact_risk = a * b / ((a + b) * (a + b + 1) * (a + b + ep['n_meas']))
    np.testing.assert_almost_equal(up.bayes_risk(ep), exact_risk, decimal=3)
    binom_ig = np.array([0.104002, 0.189223, 0.261496, 0.324283, 0.379815, 0.429613, 0.474764, 0.516069, 0.554138,
                         0.589446])
    np.testing.assert_almost_equal(up.expected_information_gain(ep), binom_ig, decimal=2)


def test_coin_model_updates_match_the_oracle(qb, oracle):
    """CoinModel (pr0 = p) through the fused update kernel, alone and under BinomialModel, vs the NumPy oracle."""
    rs = np.random.RandomState(3)
    x = rs.beta(2.0, 2.0, size=(5000, 1))
    for wrap in (False, True):
        ups = []
        for mod in (qb, oracle):
            m = mod.BinomialModel(mod.CoinModel()) if wrap else mod.CoinModel()
            np.random.seed(1)
            up = mod.SMCUpdater(m, 5000, cases.FixedPrior(x), resample_thresh=0.0)
            if wrap:
                ep = np.array([(7,), (3,), (12,)], dtype=m.expparams_dtype)
                for k, o in enumerate([2, 3, 5]):
                    up.update(o, ep[k:k + 1])
            else:
                ep = np.empty((1,), dtype=m.expparams_dtype)
                for o in [0, 1, 1, 0, 1]:
                    up.update(o, ep)
            ups.append(up)
        g, o = ups
        np.testing.assert_allclose(g.particle_weights, o.particle_weights, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(g.normalization_record, np.ravel(o.normalization_record), rtol=1e-12)
        np.testing.assert_allclose(g.est_mean(), o.est_mean(), rtol=1e-12)
        assert np.array_equal(qb.CoinModel().are_models_valid(np.array([[-0.1], [0.0], [0.5], [1.0], [1.1]])),
                              np.array([False, True, True, True, False]))



# ---------------------------------------------------------------------------
# Merge draw: sorted uniforms from exponential spacings + streaming merge with the CDF
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("kind,n", [("prec", 9), ("prec", 1000), ("prec", 4096), ("prec", 100003), ("rb", 65537),
                                    ("rb_il", 20001), ("prec", 2 ** 21), ("prec", 10 ** 7)])
def test_merge_draw_parents_and_particles_are_exact(qb, oracle, kind, n):
    """qb_lw_merge_move with its debug outputs: the uniforms are ascending, in [0, 1) and uniformly spread; every
    slot's parent is exactly min(searchsorted(cdf, u, 'right'), n - 1) on the device's own CDF (the forward walk
    equals the bisection); and the new particle is a * x[parent] + (1 - a) * mean + S @ eps with eps from the
    normal stream, bit for bit (resamplers.py:318-332 arithmetic)."""
    import torch
    model, x, w = _fused_case(qb, kind, n, 23)
    d = x.shape[1]
    seed = 4242
    res = qb.LiuWestResampler(a=0.95, rng='philox', seed=seed, scan='fast')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    mean, cov = up.est_mean(), up.est_covariance_mtx()
    S = np.real(res.h * oracle.sqrtm_psd(cov)[0])
    cdf = cloud.cdf(qb._lib.QB_SCAN_FAST_GUIDE).cpu().numpy().copy()
    u_out = torch.empty((n,), dtype=torch.float64, device=cloud.device)
    js_out = torch.empty((n,), dtype=torch.int64, device=cloud.device)
    off_e, off_n = 0, (n + 2) // 2
    seed_n = seed ^ 0x9E3779B97F4A7C15
    cloud.lw_merge_move(mean, S, 0.95, seed, off_e, seed_n, off_n, n, True, u_out=u_out, js_out=js_out)
    n_invalid, overflow = cloud.read_counter()
    u, js, got = u_out.cpu().numpy(), js_out.cpu().numpy(), cloud.x_alt.cpu().numpy().copy()
    assert np.all(np.diff(u) >= 0) and u[0] >= 0 and u[-1] < 1
    if n >= 1000:                                      # order statistics of n uniforms: KS distance ~ 1/sqrt(n)
        assert np.max(np.abs(u - (np.arange(n) + 0.5) / n)) < 4.0 / np.sqrt(n)
    want_js = np.minimum(cdf.searchsorted(u, side='right'), n - 1)
    assert np.array_equal(js, want_js)
    assert overflow == int(np.sum(cdf.searchsorted(u, side='right') >= n))
    eps = torch.empty((d * n,), dtype=torch.float64, device=cloud.device)
    cloud.rng_normal(eps, d * n, seed_n, off_n)
    e = eps.cpu().numpy().reshape(d, n)
    want = (0.95 * x[js] + (1 - 0.95) * mean)
    if d == 1:
        want = want + np.dot(S, e).T                    # one product, one sum: bit-exact
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want + np.dot(S, e).T, rtol=1e-13, atol=1e-15)
    valid = model.are_models_valid(got)
    flags = cloud._invalid[:n].cpu().numpy().astype(bool)
    assert np.array_equal(flags, ~valid) and n_invalid == int(np.sum(~valid))
    if n_invalid:
        par = cloud._parent_inv[:n].cpu().numpy()
        assert np.array_equal(par[flags], js[flags])


@pytest.mark.parametrize("kind,n", [("prec_minfreq", 50001), ("rb", 2 ** 18)])
def test_merge_mode_resample_through_the_plugin(qb, kind, n):
    """LiuWestResampler(rng='philox', scan='fast', draw='merge') through the resampler call: retries leave only
    valid particles, mean and covariance are preserved as Liu-West promises."""
    model, x, w = _fused_case(qb, kind, n, 31)
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=5, scan='fast', draw='merge')
    assert res._draw == 'merge'
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    m0, c0 = up.est_mean(), up.est_covariance_mtx()
    ess0 = up.n_ess
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        up.resample()
    assert res.last_n_iters > 1
    locs = up.particle_locations
    assert model.are_models_valid(locs).all()
    m1, c1 = up.est_mean(), up.est_covariance_mtx()
    sig = np.sqrt(np.diag(c0))
    assert np.all(np.abs(m1 - m0) < 6 * sig / np.sqrt(ess0) + 0.02 * sig)     # postselection shifts the mean a little
    assert np.all(np.abs(np.diag(c1) / np.diag(c0) - 1) < 0.15)
    assert np.all(up.particle_weights == 1.0 / n)



# ---------------------------------------------------------------------------
# f4 — MLEModel: the likelihood raised to a power inside every kernel that evaluates it
# ---------------------------------------------------------------------------
def test_mle_model_against_the_reference(qb, golden):
    """derived_models.py:681-703 through qb_likelihood and the fused update (golden vectors of the reference).
    pow() is within an ulp of glibc's; 1e-12 relative with the absolute floors of T1."""
    g = golden("mle_vectors")
    got = cases.mle_vectors(gpu_namespace(qb))
    np.testing.assert_allclose(got["prec_L"], g["prec_L"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(got["binrb_L"], g["binrb_L"], rtol=5e-12, atol=1e-300)
    np.testing.assert_allclose(got["traj_norm"], g["traj_norm"], rtol=1e-12)
    np.testing.assert_allclose(got["traj_w"], g["traj_w"], rtol=1e-11, atol=1e-15 * g["traj_w"].max())
    np.testing.assert_allclose(got["traj_mean"], g["traj_mean"], rtol=1e-12)
    report("f4_mle_traj_w_rel", relerr(got["traj_w"], g["traj_w"], floor=1e-15 * g["traj_w"].max()))
    # the power is part of the model descriptor: plain models are untouched
    assert qb.describe_model(qb.MLEModel(qb.SimplePrecessionModel(), 2.0)).likelihood_power == 2.0
    assert qb.describe_model(qb.SimplePrecessionModel()).likelihood_power == 1.0


# ---------------------------------------------------------------------------
# f4 — time-dependent and noisy decorators on the device (derived_models.py:148-220, 705-963;
# tomography/models.py:228-272) against golden trajectories of the UNMODIFIED reference
# ---------------------------------------------------------------------------
def _f4_check(got, g, key, x_rtol=1e-9):
    """Free-running trajectories incl. resampling under the legacy seed: same resample count, records 1e-9, particles
    1e-9 (the device evaluates cos^2 to within an ulp of NumPy's; the walk / noise arithmetic is operation for
    operation the reference's)."""
    assert int(got[key + '_rc']) == int(g[key + '_rc'])
    np.testing.assert_allclose(got[key + '_norm'], g[key + '_norm'], rtol=1e-9)
    np.testing.assert_allclose(got[key + '_x'], g[key + '_x'], rtol=x_rtol, atol=1e-13)
    np.testing.assert_allclose(got[key + '_w'], g[key + '_w'], rtol=1e-7, atol=1e-16)
    report("f4_%s_x_maxabs" % key, float(np.max(np.abs(got[key + '_x'] - g[key + '_x']))))


def test_f4_random_walk_and_poisoned_models_against_the_reference(qb, golden):
    g = golden("random_walk_vectors")
    got = cases.random_walk_vectors(gpu_namespace(qb))
    for key in ("rw", "grw_fixed", "grw_learn", "ale", "mle"):
        _f4_check(got, g, key)
    assert np.array_equal(np.asarray(got['grw_learn_valid'], dtype=bool), g['grw_learn_valid'])


def test_f4_diffusive_tomography_against_the_reference(qb, golden):
    g = golden("diffusive_vectors")
    got = cases.diffusive_vectors(gpu_namespace(qb))
    assert np.array_equal(np.asarray(got['valid'], dtype=bool), g['valid'])
    assert int(got['rc']) == int(g['rc'])
    np.testing.assert_allclose(got['norm'], g['norm'], rtol=1e-8)
    # canonicalisation: Hermitian Jacobi on the device vs np.linalg.eig in the reference agree to ~1e-10 per call
    # (test_t6); the state is re-projected after each of the 24 updates and 3 resamples of this run, and states
    # clipped at the boundary of the Bloch ball (an eigenvalue ~ 0) are the most sensitive: measured 9e-9 worst case
    np.testing.assert_allclose(got['x'], g['x'], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(np.average(got['x'], axis=0, weights=got['w']),
                               np.average(g['x'], axis=0, weights=g['w']), rtol=1e-6, atol=1e-9)   # north_star
    np.testing.assert_allclose(got['w'], g['w'], rtol=1e-6, atol=1e-15)
    report("f4_diffusive_x_maxabs", float(np.max(np.abs(got['x'] - g['x']))))


@pytest.mark.parametrize("rng", ["mt19937", "philox"])
def test_f4_device_noise_sources(qb, rng):
    """The decorators' noise from the device generators: 'mt19937' continues the legacy stream (same trajectory as
    the host-drawn one up to the 1-ulp normals), 'philox' is statistically equivalent (random-walk variance grows as
    sigma^2 per update)."""
    n, sigma = 200000, 3e-3
    rs = np.random.RandomState(4)
    prior = 0.5 + 0.01 * rs.randn(n, 1)
    model = qb.GaussianRandomWalkModel(qb.SimplePrecessionModel(), fixed_covariance=np.array([sigma ** 2]))
    res = qb.LiuWestResampler(a=0.98, rng=rng, seed=9, scan='exact' if rng == 'mt19937' else 'fast')
    np.random.seed(3)
    up = qb.SMCUpdater(model, n, cases.FixedPrior(prior), resampler=res, resample_thresh=0.0)
    v0 = float(np.var(up.particle_locations))
    for k in range(8):
        up.update(k % 2, np.array([1e-3]))                # (likelihood ~ flat: the cloud only diffuses)
    v1 = float(np.var(up.particle_locations))
    assert abs((v1 - v0) / (8 * sigma ** 2) - 1) < 0.03
    if rng == 'mt19937':
        np.random.seed(3)
        ref = qb.SMCUpdater(model, n, cases.FixedPrior(prior), resample_thresh=0.0)
        for k in range(8):
            ref.update(k % 2, np.array([1e-3]))
        np.testing.assert_allclose(up.particle_locations, ref.particle_locations, rtol=1e-12, atol=1e-15)


def test_f4_unsupported_decorator_variants_fail_loudly(qb):
    with pytest.raises(qb.UnsupportedModelError):
        qb.GaussianRandomWalkModel(qb.SimplePrecessionModel(), diagonal=False)
    with pytest.raises(qb.UnsupportedModelError):
        qb.GaussianRandomWalkModel(qb.SimplePrecessionModel(), fixed_covariance=np.array([1e-6]),
                                   model_transformation=(np.log, np.exp))


def test_design_under_mle_model_follows_the_reference_complement(qb, golden):
    """with a likelihood power the device must form the last outcome as 1 - sum(others ** gamma) like
    smc.py:589, not evaluate L_last ** gamma."""
    g = golden("mle_design_vectors")
    got = cases.mle_design_vectors(gpu_namespace(qb))
    for key in ("g2_risk", "g2_ig", "g05_risk", "g05_ig"):
        ok = np.isfinite(g[key])
        assert ok.any()
        np.testing.assert_allclose(got[key][ok], g[key][ok], rtol=1e-9, err_msg=key)


@pytest.mark.parametrize("binomial", [False, True])
def test_fast_math_likelihoods_within_1e12_of_the_strict_path(qb, binomial):
    """SMCUpdater(fast_math=True): p ** m and the binomial pmf by integer powers (qb_model.fast_math) against the
    reference's operation sequence (pow, exp(logC + k log p + (n-k) log1p(-p))): weights within 4e-13 relative per update,
    records and n_ess likewise, over sequence lengths up to 800 and every count k of n_meas = 25."""
    rs = np.random.RandomState(6)
    n = 200000
    x = np.column_stack([0.8 + 0.2 * rs.random_sample(n), 0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    model = qb.RandomizedBenchmarkingModel()
    if binomial:
        model = qb.BinomialModel(model)
    eps = np.empty((26,), dtype=model.expparams_dtype)
    eps['m'] = np.linspace(1, 800, 26).astype(int)
    if binomial:
        eps['n_meas'] = 25
    outs = {}
    for fast in (False, True):
        up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resample_thresh=0.0, fast_math=fast)
        ws = []
        for k in range(26):
            up.update(k if binomial else k % 2, eps[k:k + 1])
            if k in (0, 5, 25):
                ws.append(up.particle_weights.copy())
        outs[fast] = (ws, np.array(up.normalization_record), up.n_ess)
    (w0, r0, e0), (w1, r1, e1) = outs[False], outs[True]
    for steps, a, b in zip((1, 6, 26), w0, w1):
        rel = np.abs(b - a) / np.maximum(np.abs(a), 1e-300)
        worst = int(np.argmax(rel))
        # (<= 4e-13 per update: a few ulp of p ** m times the conditioning (n - k) pr0 / (1 - pr0) of the pmf, and the
        # strict path's exp() of an argument near -100; the weights are products, so it adds up over the updates)
        assert rel[worst] <= steps * 4e-13, (steps, rel[worst], a[worst], b[worst], x[worst])
    np.testing.assert_allclose(r1, r0, rtol=1e-12)
    assert abs(e1 - e0) <= 1e-11 * e0
    report("fast_math_%s_weights_rel" % ("binom" if binomial else "rb"), relerr(w1[-1], w0[-1], floor=1e-300))
