"""GPU tests of the binned multinomial Liu-West resample (csrc/qb_binned.cu; resamplers.py:266-273, 308-372, 390-392).

The binned draw is the default of ``LiuWestResampler(rng='philox', scan='fast')``.  It is not index-identical to the
reference (that is the parity mode's job: ``rng='numpy'|'mt19937', scan='exact'``); what it promises, and what is
checked here through the C ABI, is
  * the weighted moments of its first pass equal the reference's formulas (distributions.py:337-399) to 1e-12;
  * the per-bin offspring counts are EXACTLY the histogram of the Philox uniforms over the bin-level CDF, the output
    offsets their prefix sum;
  * every slot's parent lies in the slot's bin and is the right-bisection of the bin-local CDF at the slot's uniform;
  * every new particle is a * x[parent] + (1 - a) * mean + S @ eps with eps from the normal stream — bit for bit for
    d = 1 (resamplers.py:325,332 arithmetic, one rounding per ufunc);
  * the invalid list, the retry loop, the fused uniform weights (resamplers.py:390-392) and their stats block;
  * the LAW of the draw: offspring counts per parent are multinomial in the normalised weights (chi-square).
"""
import warnings

import numpy as np
import pytest

import cases
from test_gpu_parity import _fused_case

pytestmark = pytest.mark.gpu

BIN = 2048
SEG = 2 * BIN
GOLD = 0x9E3779B97F4A7C15


@pytest.fixture(scope="module")
def qb():
    import qinfer_b200
    return qinfer_b200


@pytest.fixture(scope="module")
def oracle():
    import smc_oracle
    return smc_oracle


def _align256(b):
    return (b + 255) // 256 * 256


def _read_ws(cloud):
    """bounds (T+1), counts (T), offs (T+1), nseg from the binned workspace (layout of bin_layout(), qb_binned.cu)."""
    import torch
    torch.cuda.synchronize()
    T = (cloud.n + BIN - 1) // BIN
    raw = cloud._bin_ws.cpu().numpy().view(np.uint8)
    o_bounds = 512
    o_counts = o_bounds + _align256((T + 1) * 8)
    o_offs = o_counts + _align256((T + 2) * 4)
    bounds = raw[o_bounds:o_bounds + (T + 1) * 8].view(np.float64).copy()
    counts = raw[o_counts:o_counts + T * 4].view(np.uint32).astype(np.int64)
    offs = raw[o_offs:o_offs + (T + 1) * 8].view(np.int64).copy()
    nseg = int(raw[12:16].view(np.uint32)[0])
    return T, bounds, counts, offs, nseg


def _stream(cloud, n, seed, off, normal=False):
    import torch
    out = torch.empty((n,), dtype=torch.float64, device=cloud.device)
    (cloud.rng_normal if normal else cloud.rng_uniform)(out, n, seed, off)
    return out.cpu().numpy()


def _weights(kind, n, rs):
    if kind == "sorted":          # smooth in the index: almost all mass in the first bins (heavy bins -> many segments)
        w = np.exp(-np.arange(n) / (n / 40.0))
    elif kind == "zero_bins":     # whole bins of zero weight, a bin with a single non-zero particle
        w = rs.random_sample(n)
        w[BIN:3 * BIN] = 0.0
        if n > 5 * BIN:
            w[4 * BIN:5 * BIN] = 0.0
            w[4 * BIN + 17] = 0.3
    else:
        w = rs.random_sample(n) ** 4
    return w / w.sum()


@pytest.mark.parametrize("kind,n,n_new,wkind", [
    ("prec", 9, None, "rand"), ("prec", 2048, None, "rand"), ("prec", 2049, None, "rand"),
    ("prec", 100003, None, "rand"), ("prec", 100003, 250001, "sorted"), ("prec", 6 * BIN + 5, 40000, "zero_bins"),
    ("prec_minfreq", 50001, None, "case"), ("rb", 65537, None, "case"), ("rb", 30000, 29999, "case"),
    ("rb_il", 20001, None, "case"), ("prec", 2 ** 21, None, "case"), ("prec", 10 ** 7, None, "case")])
def test_binned_passes_are_exact(qb, oracle, kind, n, n_new, wkind):
    import torch
    model, x, w = _fused_case(qb, kind, n, 23)
    if wkind != "case":
        w = _weights(wkind, n, np.random.RandomState(5))
    d = x.shape[1]
    n_new = n if n_new is None else n_new
    seed, a = 4242, 0.95
    res = qb.LiuWestResampler(a=a, rng='philox', seed=seed, scan='fast')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    cloud.count_mode = qb._lib.QB_COUNT_HISTOGRAM          # (the counts are then checked against the uniforms themselves)
    off_u, off_v, off_n = 3, 3 + (n_new + 1) // 2, 3 + 2 * ((n_new + 1) // 2)

    # ---- pass 1 + 2 -------------------------------------------------------------------------------------------
    tag = cloud.binned_prepare(n_new, seed, off_u)
    s0, mean, m2 = cloud.binned_moments_wait(tag)
    torch.cuda.synchronize()
    assert abs(s0 - 1.0) < 1e-12
    np.testing.assert_allclose(mean, np.dot(w, x), rtol=1e-12, atol=1e-15)                 # distributions.py:337-348
    np.testing.assert_allclose(m2, np.einsum('i,im,in->mn', w, x, x), rtol=1e-12, atol=1e-15)   # :386-387
    dev = cloud.moments_out.cpu().numpy()
    assert dev[0] == s0 and np.array_equal(dev[1:1 + d], mean) and np.array_equal(dev[1 + d:].reshape(d, d), m2)
    T, bounds, counts, offs, nseg = _read_ws(cloud)
    pad = np.zeros(T * BIN)
    pad[:n] = w
    want_bounds = np.concatenate([[0.0], np.cumsum(pad.reshape(T, BIN).sum(axis=1))])
    np.testing.assert_allclose(bounds, want_bounds, rtol=0, atol=1e-12)
    u = _stream(cloud, n_new, seed, off_u)
    bins = np.clip(np.searchsorted(bounds[:T], u * bounds[T], side='right') - 1, 0, T - 1)
    assert np.array_equal(counts, np.bincount(bins, minlength=T))           # the multinomial counts, exactly
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(counts)])) and offs[T] == n_new
    assert nseg == int(np.sum((counts + SEG - 1) // SEG))

    # ---- pass 3 -------------------------------------------------------------------------------------------------
    cov = m2 - np.outer(mean, mean)
    S = np.real(res.h * oracle.sqrtm_psd(cov)[0])
    js_out = torch.empty((n_new,), dtype=torch.int64, device=cloud.device)
    fuse = n_new == n
    tag = cloud.binned_move(mean, S, a, seed, off_v, seed ^ GOLD, off_n, n_new, True, fuse_weights=fuse, js_out=js_out)
    n_invalid, n_clamped, drawn = cloud.binned_counters_wait(tag)
    torch.cuda.synchronize()
    assert drawn == n_new
    js, got = js_out.cpu().numpy(), cloud.x_alt.cpu().numpy().copy()
    slot_bin = np.searchsorted(offs, np.arange(n_new), side='right') - 1   # slot i belongs to the bin whose range holds it
    assert np.array_equal(js // BIN, slot_bin)
    # the parent is the right-bisection of the bin-local CDF at v = u2 * (bin total); the device scans the bin in
    # another association order than np.cumsum, so compare with a tolerance of a few ulp of the bin total
    u2 = _stream(cloud, n_new, seed, off_v)
    local = np.cumsum(pad.reshape(T, BIN), axis=1)
    top = local[slot_bin, -1]
    v = u2 * top
    j = js - slot_bin * BIN
    tol = 1e-12 * top
    upper = local[slot_bin, j]
    lower = np.where(j > 0, local[slot_bin, np.maximum(j - 1, 0)], 0.0)
    last = np.minimum((slot_bin + 1) * BIN, n) - slot_bin * BIN - 1
    assert np.all(lower <= v + tol)
    assert np.all((v < upper + tol) | (j == last))                          # (the clamp of distributions.py:330-333)
    assert n_clamped <= 2
    if n_clamped == 0:
        assert np.all(w[js] > 0)                                            # a zero-weight particle is never drawn
    e = _stream(cloud, d * n_new, seed ^ GOLD, off_n, normal=True).reshape(d, n_new)
    want = a * x[js] + (1 - a) * mean                                       # resamplers.py:325
    if d == 1:
        want = want + np.dot(S, e).T                                        # resamplers.py:332
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want + np.dot(S, e).T, rtol=1e-13, atol=1e-15)
    bad = ~np.asarray(model.are_models_valid(got), dtype=bool)
    assert n_invalid == int(bad.sum())
    lst = cloud._bin_list[:n_invalid].cpu().numpy()
    assert np.array_equal(np.sort(lst & 0xFFFFFFFF), np.nonzero(bad)[0])
    assert np.array_equal(js[lst & 0xFFFFFFFF], lst >> 32)
    if fuse:
        assert np.all(cloud.w_alt.cpu().numpy() == 1.0 / n)                 # resamplers.py:390-392
        st = cloud.stats_alt.cpu().numpy()
        assert st[0] == 1.0 and st[1] == 1.0 / n and st[4] == 1.0 and st[5] == float(n) and st[8] == 0.0
    if kind in ("prec_minfreq", "rb", "rb_il"):
        assert n_invalid > 0
        # one retry round: every listed slot is re-centred on the parent of a uniformly random slot (the law of the
        # reference's mus[:k] re-slicing, resamplers.py:372) with normals indexed by slot
        off_r = off_n + (d * n_new + 1) // 2
        tag = cloud.binned_retry(mean, S, a, seed ^ GOLD, off_r, n_new, 1, seed_v=seed)
        left, used, listed = cloud.binned_retry_wait(tag)
        assert used == 1 and listed == n_invalid
        torch.cuda.synchronize()
        got2 = cloud.x_alt.cpu().numpy()
        e2 = _stream(cloud, d * n_new, seed ^ GOLD, off_r, normal=True).reshape(d, n_new)
        slots = np.nonzero(bad)[0]
        uq = _stream(cloud, n_new, seed, off_r)
        donor = np.minimum((uq[slots] * n_new).astype(np.int64), n_new - 1)
        assert np.array_equal(cloud._bin_parents[:n_new].cpu().numpy(), js)
        want2 = a * x[js[donor]] + (1 - a) * mean + np.dot(S, e2[:, slots]).T
        np.testing.assert_allclose(got2[slots], want2, rtol=1e-13, atol=1e-15)
        assert np.array_equal(got2[~bad], got[~bad])
        still = ~np.asarray(model.are_models_valid(got2[slots]), dtype=bool)
        assert left == int(still.sum())
        lst2 = cloud._bin_list[:n_invalid].cpu().numpy()
        valid2 = np.asarray(model.are_models_valid(got2), dtype=bool)
        resolved = valid2[lst & 0xFFFFFFFF]
        assert np.array_equal(lst2 < 0, resolved) and np.array_equal(lst2[~resolved], lst[~resolved])


@pytest.mark.parametrize("wkind", ["rand", "sorted", "zero_bins"])
def test_binned_draw_is_multinomial_in_the_weights(qb, oracle, wkind):
    """Offspring counts per parent against n_new * w: Pearson chi-square over the parents with an expected count of at
    least 5 (the others pooled), within 5 sigma of its degrees of freedom; and the bin-level counts likewise."""
    import torch
    n, n_new = 6 * BIN + 5, 3 * 10 ** 6
    rs = np.random.RandomState(3)
    x = rs.random_sample((n, 1))
    w = _weights(wkind, n, rs)
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=99, scan='fast')
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    tag = cloud.binned_prepare(n_new, 99, 0)
    _, mean, m2 = cloud.binned_moments_wait(tag)
    js_out = torch.empty((n_new,), dtype=torch.int64, device=cloud.device)
    dst = torch.empty((n_new, 1), dtype=torch.float64, device=cloud.device)
    tag = cloud.binned_move(mean, np.zeros((1, 1)), 0.98, 99, (n_new + 1) // 2, 5, 0, n_new, False, dst=dst, js_out=js_out)
    assert cloud.binned_counters_wait(tag)[2] == n_new
    got = np.bincount(js_out.cpu().numpy(), minlength=n).astype(float)
    expect = n_new * w
    big = expect >= 5
    chi2 = float(np.sum((got[big] - expect[big]) ** 2 / expect[big]))
    dof = int(big.sum())
    rest_e, rest_g = expect[~big].sum(), got[~big].sum()
    if rest_e > 5:
        chi2 += (rest_g - rest_e) ** 2 / rest_e
        dof += 1
    assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof), (chi2, dof)
    assert np.all(got[w == 0] == 0)


def test_binned_resample_is_deterministic_and_seeded(qb):
    model, x, w = _fused_case(qb, "prec_minfreq", 50001, 31)
    outs = []
    for seed in (5, 5, 6):
        res = qb.LiuWestResampler(a=0.98, rng='philox', seed=seed, scan='fast', draw='binned')
        up = qb.SMCUpdater(model, 50001, cases.FixedPrior(x), resampler=res)
        up.particle_weights = w
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            up.resample()
        assert res.last_n_iters > 1
        outs.append(up.particle_locations.copy())
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("kind,n", [("prec", 1000), ("prec_minfreq", 50001), ("rb", 2 ** 18), ("rb_il", 20001)])
def test_binned_resample_through_the_plugin(qb, kind, n):
    """The default device-RNG draw through SMCUpdater.resample(): only valid particles remain, uniform weights,
    n_ess = n, mean and covariance preserved as Liu-West promises (resamplers.py:206-221)."""
    model, x, w = _fused_case(qb, kind, n, 31)
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=5, scan='fast')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    m0, c0 = up.est_mean(), up.est_covariance_mtx()
    ess0 = up.n_ess
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        up.resample()
    if kind != "prec":
        assert res.last_n_iters > 1
    locs = up.particle_locations
    assert locs.shape == x.shape and np.asarray(model.are_models_valid(locs)).all()
    assert np.all(up.particle_weights == 1.0 / n) and up.n_ess == pytest.approx(n)
    m1, c1 = up.est_mean(), up.est_covariance_mtx()
    sig = np.sqrt(np.diag(c0))
    assert np.all(np.abs(m1 - m0) < 6 * sig / np.sqrt(ess0) + 0.02 * sig)     # postselection shifts the mean a little
    assert np.all(np.abs(np.diag(c1) / np.diag(c0) - 1) < 0.15)
    # and the updater keeps working on the swapped buffers
    if kind.startswith("prec"):
        up.update(1, np.array([0.7]))
        assert np.isfinite(up.n_ess) and up.n_ess <= n


def test_binned_and_guided_trajectories_agree_statistically(qb):
    """A C2-shaped run (exp-sparse schedule, resampling on) with the binned and with the guided draw: same number of
    resamples +-1 and posterior means within 5 standard errors of each other."""
    n, steps = 200000, 60
    rs = np.random.RandomState(2)
    prior = rs.random_sample((n, 1))
    ts = (9.0 / 8.0) ** np.arange(steps)
    outcomes = (rs.random_sample(steps) >= np.cos(ts * 0.5 / 2) ** 2).astype(int)
    out = []
    for draw in ("binned", "guided"):
        res = qb.LiuWestResampler(a=0.98, rng='philox', seed=11, scan='fast', draw=draw)
        up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(prior), resampler=res, lazy=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k in range(steps):
                up.update(int(outcomes[k]), ts[k:k + 1])
        out.append((up.resample_count, up.est_mean()[0], up.est_covariance_mtx()[0, 0], up.n_ess))
    (rb_, mb, cb, eb), (rg, mg, cg, eg) = out
    assert abs(rb_ - rg) <= 1 and rb_ >= 3
    se = np.sqrt(cb / eb + cg / eg)
    assert abs(mb - mg) < 5 * se + 1e-9
    assert abs(mb - 0.5) < 0.01


@pytest.mark.parametrize("kind,n", [("prec", 5000), ("rb", 40001), ("rb_il", 20001)])
def test_binned_device_constants_match_the_host_formulas(qb, oracle, kind, n):
    """qb_lw_binned_resample derives S = h * sqrtm_psd(cov) and (1 - a) * mean in its first kernel (resamplers.py:
    266-305, utils.py:593-607): bit-identical to the host path for d = 1, 1e-13 of ||S|| for d = 3, 4 (Jacobi vs
    LAPACK); and the particles it moves with them equal the host-constant path's to the same tolerance."""
    import torch
    model, x, w = _fused_case(qb, kind, n, 41)
    d = x.shape[1]
    a = 0.9
    res = qb.LiuWestResampler(a=a, rng='philox', seed=77, scan='fast')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    off_u, off_v, off_n = 0, (n + 1) // 2, 2 * ((n + 1) // 2)
    tag = cloud.binned_resample(n, a, res.h, 1e-10, 77, off_u, off_v, 77 ^ GOLD, off_n, False, 0, False)
    _, mean, m2 = cloud.binned_moments_wait(tag)
    assert cloud.binned_flags()[0] == 0
    cloud.binned_counters_wait(tag)
    torch.cuda.synchronize()
    got_dev = cloud.x_alt.cpu().numpy().copy()
    raw = cloud._bin_ws.cpu().numpy().view(np.uint8)
    consts = raw[128:128 + 20 * 8].view(np.float64)
    cov = m2 - np.outer(mean, mean)
    S0, err = oracle.sqrtm_psd(cov)
    S = np.real(res.h * S0)
    ms = (1 - a) * mean
    if d == 1:
        assert consts[0] == S[0, 0] and consts[16] == ms[0]
    else:
        np.testing.assert_allclose(consts[:d * d].reshape(d, d), S, rtol=0, atol=1e-13 * np.linalg.norm(S))
        assert np.array_equal(consts[16:16 + d], ms)
    assert abs(cloud.binned_flags()[1] - err) < 1e-12 * max(np.linalg.norm(cov), 1e-300) + 1e-18
    # same streams, host-supplied constants
    tag = cloud.binned_prepare(n, 77, off_u)
    cloud.binned_moments_wait(tag)
    tag = cloud.binned_move(mean, S, a, 77, off_v, 77 ^ GOLD, off_n, n, False)
    cloud.binned_counters_wait(tag)
    torch.cuda.synchronize()
    got_host = cloud.x_alt.cpu().numpy()
    if d == 1:
        assert np.array_equal(got_dev, got_host)
    else:
        np.testing.assert_allclose(got_dev, got_host, rtol=0, atol=1e-12)


def test_binned_zero_covariance_warns_and_uses_the_small_covariance(qb):
    """resamplers.py:288-293: a cloud collapsed on one point -> ResamplerWarning, cov = zero_cov_comp * I."""
    n = 8192                              # 1/n, the sums and x are exact in binary: the covariance is exactly 0
    x = np.full((n, 1), 0.5)
    res = qb.LiuWestResampler(a=0.98, rng='philox', seed=1, scan='fast', zero_cov_comp=1e-6)
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x), resampler=res)
    with pytest.warns(qb.ResamplerWarning, match="zero norm"):
        up.resample()
    locs = up.particle_locations
    assert abs(locs.mean() - 0.5) < 1e-4
    assert abs(locs.std() / (res.h * 1e-3) - 1) < 0.05                      # S = h * sqrt(1e-6)


def test_binned_retry_own_parent_variant(qb, oracle):
    """own_mean=True: a still-invalid particle is retried around ITS OWN parent (textbook Liu-West)."""
    import torch
    model, x, w = _fused_case(qb, "prec_minfreq", 50001, 23)
    n, seed, a = 50001, 4242, 0.95
    res = qb.LiuWestResampler(a=a, rng='philox', seed=seed, scan='fast')
    up = qb.SMCUpdater(model, n, cases.FixedPrior(x), resampler=res)
    up.particle_weights = w
    cloud = up._cloud
    tag = cloud.binned_prepare(n, seed, 0)
    _, mean, m2 = cloud.binned_moments_wait(tag)
    S = np.real(res.h * oracle.sqrtm_psd(m2 - np.outer(mean, mean))[0])
    js_out = torch.empty((n,), dtype=torch.int64, device=cloud.device)
    off_n = 2 * ((n + 1) // 2)
    tag = cloud.binned_move(mean, S, a, seed, (n + 1) // 2, seed ^ GOLD, off_n, n, True, js_out=js_out)
    n_invalid = cloud.binned_counters_wait(tag)[0]
    torch.cuda.synchronize()
    got = cloud.x_alt.cpu().numpy().copy()
    bad = ~np.asarray(model.are_models_valid(got), dtype=bool)
    assert n_invalid == bad.sum() > 0
    off_r = off_n + (n + 1) // 2
    tag = cloud.binned_retry(mean, S, a, seed ^ GOLD, off_r, n, 1, own_mean=True)
    cloud.binned_retry_wait(tag)
    torch.cuda.synchronize()
    e2 = _stream(cloud, n, seed ^ GOLD, off_r, normal=True)
    slots = np.nonzero(bad)[0]
    want = a * x[js_out.cpu().numpy()[slots]] + (1 - a) * mean + (S[0, 0] * e2[slots])[:, None]
    assert np.array_equal(cloud.x_alt.cpu().numpy()[slots], want)


@pytest.mark.parametrize("n,p", [(5, 0.3), (40, 0.5), (1000, 0.01), (1000, 0.2), (10 ** 7, 0.5), (10 ** 7, 1e-6),
                                 (10 ** 7, 1.0 - 3e-6), (12345678, 0.0372), (2 ** 31 - 1, 0.37), (3000, 0.9871)])
def test_device_binomial_sampler_is_exact(qb, n, p):
    """The sampler behind the binomial-tree counts (inversion below a mean of 30, BTPE above) against
    scipy.stats.binom: Pearson chi-square of 4e5 variates over the outcomes with an expected count >= 10 (tails
    pooled), within 5 sigma of its degrees of freedom; mean and variance within 5 standard errors."""
    import scipy.stats
    import torch
    from qinfer_b200.engine import _ptr, _stream
    lib = qb._lib.load()
    count = 400000
    out = torch.empty((count,), dtype=torch.int64, device='cuda')
    qb._lib.check(lib.qb_binomial_sample(n, p, count, 77, 1000, _ptr(out), _stream()))
    v = out.cpu().numpy()
    assert v.min() >= 0 and v.max() <= n
    mean, var = n * p, n * p * (1 - p)
    assert abs(v.mean() - mean) < 5 * np.sqrt(var / count) + 1e-12
    assert abs(v.var() - var) < 5 * var * np.sqrt(2.0 / count + (1 / max(var, 1e-300) - 6 / n) / count) + 1e-12
    lo, hi = int(max(0, np.floor(mean - 8 * np.sqrt(var) - 10))), int(min(n, np.ceil(mean + 8 * np.sqrt(var) + 10)))
    ks = np.arange(lo, hi + 1)
    expect = count * scipy.stats.binom.pmf(ks, n, p)
    got = np.bincount(np.clip(v - lo, 0, hi - lo), minlength=ks.size).astype(float)
    big = expect >= 10
    chi2 = float(np.sum((got[big] - expect[big]) ** 2 / expect[big]))
    dof = int(big.sum())
    rest_e, rest_g = count - expect[big].sum(), count - got[big].sum()
    if rest_e > 10:
        chi2 += (rest_g - rest_e) ** 2 / rest_e
        dof += 1
    assert abs(chi2 - (dof - 1)) < 5 * np.sqrt(2 * dof) + 5, (chi2, dof)


@pytest.mark.parametrize("wkind,n,n_new", [("rand", 6 * BIN + 5, 3 * 10 ** 6), ("sorted", 100003, 250001),
                                           ("zero_bins", 6 * BIN + 5, 40000), ("rand", 10 ** 7, 10 ** 7),
                                           ("rand", 9, 1000), ("rand", 2 ** 23 + 77, 2 ** 23 + 77)])
def test_tree_counts_are_multinomial_in_the_bin_masses(qb, wkind, n, n_new):
    """QB_COUNT_TREE: the bin counts sum to n_new exactly, bins of zero mass get nothing, and over 12 independent
    seeds the counts are consistent with Multinomial(n_new; bin masses) (pooled Pearson chi-square, 5 sigma)."""
    rs = np.random.RandomState(2)
    x = rs.random_sample((n, 1))
    w = _weights(wkind, n, rs)
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(x))
    up.particle_weights = w
    cloud = up._cloud
    cloud.count_mode = qb._lib.QB_COUNT_TREE
    T = (n + BIN - 1) // BIN
    pad = np.zeros(T * BIN)
    pad[:n] = w
    mass = pad.reshape(T, BIN).sum(axis=1)
    chi2 = dof = 0.0
    for seed in range(12):
        tag = cloud.binned_prepare(n_new, 1000 + seed, 5 * seed)
        cloud.binned_moments_wait(tag)
        import torch
        torch.cuda.synchronize()
        _, bounds, counts, offs, nseg = _read_ws(cloud)
        assert counts.sum() == n_new and offs[T] == n_new
        assert np.array_equal(offs, np.concatenate([[0], np.cumsum(counts)]))
        assert np.all(counts[mass == 0] == 0)
        assert nseg == int(np.sum((counts + SEG - 1) // SEG))
        e = n_new * mass
        big = e >= 10
        if big.sum() > 1:
            chi2 += float(np.sum((counts[big] - e[big]) ** 2 / e[big]))
            dof += big.sum() - (1 if big.all() else 0)
    if dof > 0:
        assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof) + 5, (chi2, dof)
    # different seeds give different counts, the same seed the same
    tag = cloud.binned_prepare(n_new, 1000, 0)
    cloud.binned_moments_wait(tag)
    c0 = _read_ws(cloud)[2]
    tag = cloud.binned_prepare(n_new, 1000, 0)
    cloud.binned_moments_wait(tag)
    assert np.array_equal(_read_ws(cloud)[2], c0)
