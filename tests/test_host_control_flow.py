"""The host half of the update path — speculation, fusion, roll-back, the zero-weight policies and the resample
trigger (qinfer_b200/smc.py) — driven on CPU against a NumPy double of the device cloud (tests/fake_cloud.py) and
compared with the oracle's plain call-by-call updater (smc.py:388-487 restated)."""
import warnings

import numpy as np
import pytest

import cases
import qinfer_b200 as qb
import smc_oracle as oracle
from fake_cloud import FakeCloud
from qinfer_b200.smc import SMCUpdater


class HostOnlyUpdater(SMCUpdater):
    """The product's updater with its device cloud swapped for the NumPy double."""

    def _rebuild_cloud(self, n):
        launches = self._cloud.launches if self._cloud is not None else 0
        self._cloud = FakeCloud(self._desc, n)
        self._cloud.launches = launches
        self._host_locs = self._host_weights = None


def _run(up, ts, outcomes, batch=None):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if batch is None:
            for t, o in zip(ts, outcomes):
                up.update(int(o), np.array([t]))
        else:
            up.batch_update(np.asarray(outcomes), np.asarray(ts), resample_interval=batch)
    return up


def _pair(n, seed, lazy, fuse, **kw):
    inp = cases.precession_inputs(n_particles=n, n_updates=80, seed=seed)
    np.random.seed(0)
    ref = oracle.SMCUpdater(oracle.SimplePrecessionModel(), n, cases.FixedPrior(inp['prior']), **kw)
    np.random.seed(0)
    # the oracle's resampler works on host arrays through particle_locations / particle_weights: the updater's
    # "foreign resampler" path (smc.py: resample()), so both sides draw the same NumPy variates
    got = HostOnlyUpdater(oracle.SimplePrecessionModel(), n, cases.FixedPrior(inp['prior']), lazy=lazy, fuse=fuse,
                          resampler=oracle.LiuWestResampler(), **kw)
    return inp, ref, got


@pytest.mark.parametrize("lazy,fuse", [(False, None), (True, 1), (True, 3), (True, 8)])
def test_speculation_and_fusion_reproduce_the_call_by_call_trajectory(lazy, fuse):
    """Same data, same seed: records, n_ess bookkeeping, resample count and posterior agree with the oracle whether
    every update is settled at once, launched speculatively one ahead, or fused 3 / 8 per launch with roll-back
    to the step that triggered a resample."""
    inp, ref, got = _pair(800, 1234, lazy, fuse)
    np.random.seed(1)
    _run(ref, inp['ts'], inp['outcomes'])
    np.random.seed(1)
    _run(got, inp['ts'], inp['outcomes'])
    # one update per launch is bit-identical to the oracle; a fused launch omits the renormalisation between its
    # steps (a few ulp), which late evolution times (t up to 1e4 here) amplify
    rtol = 1e-12 if not (fuse and fuse > 1) else 1e-6
    assert got.resample_count == ref.resample_count > 3
    np.testing.assert_allclose(got.normalization_record, np.ravel(ref.normalization_record), rtol=rtol)
    assert got.min_n_ess == pytest.approx(ref.min_n_ess, rel=rtol)
    assert got.n_ess == pytest.approx(ref.n_ess, rel=rtol)
    np.testing.assert_allclose(got.est_mean(), ref.est_mean(), rtol=rtol)
    np.testing.assert_allclose(got.particle_weights, ref.particle_weights, rtol=100 * rtol, atol=1e-18)
    log = got._cloud.log
    if lazy:
        assert any(g for _, _, g, _ in log)                               # speculative launches went out ...
        assert any(c for _, _, _, c in log)                               # ... and some cancelled themselves
    if lazy and fuse and fuse > 1:
        assert max(k for _, k, _, _ in log) == fuse                       # fused launches happened


def test_batch_update_matches_the_reference_loop():
    """smc.py:459-487: resample check every `resample_interval`-th datum only."""
    inp, ref, got = _pair(600, 77, True, 8)
    np.random.seed(2)
    _run(ref, inp['ts'], inp['outcomes'], batch=5)
    np.random.seed(2)
    _run(got, inp['ts'], inp['outcomes'], batch=5)
    assert got.resample_count == ref.resample_count > 0
    np.testing.assert_allclose(got.normalization_record, np.ravel(ref.normalization_record), rtol=1e-6)
    np.testing.assert_allclose(got.est_mean(), ref.est_mean(), rtol=1e-6)


@pytest.mark.parametrize("policy", ["skip", "warn", "ignore", "error", "reset"])
@pytest.mark.parametrize("lazy,fuse", [(False, None), (True, 4)])
def test_zero_weight_policies_under_speculation(policy, lazy, fuse):
    """smc.py:423-436: a datum that kills every weight (all particles at omega = 0, outcome 1 has likelihood 0) in the
    middle of a queue of updates."""
    n = 64
    x = np.zeros((n, 1))
    steps = [(0, 1.0), (0, 2.0), (1, 3.0), (0, 4.0), (0, 5.0)]            # the third datum is impossible

    def drive(up):
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            err = None
            try:
                for o, t in steps:
                    up.update(o, np.array([t]))
                up.n_ess                                                   # settles everything that is pending
            except RuntimeError as e:
                err = str(e)
        return err, [str(w.message) for w in wlist if "All particle weights are zero" in str(w.message)]

    kw = dict(zero_weight_policy=policy, resample_thresh=0.0)
    ref = oracle.SMCUpdater(oracle.SimplePrecessionModel(), n, cases.FixedPrior(x), **kw)
    got = HostOnlyUpdater(oracle.SimplePrecessionModel(), n, cases.FixedPrior(x), lazy=lazy, fuse=fuse,
                          resampler=oracle.LiuWestResampler(), **kw)
    err_r, warn_r = drive(ref)
    err_g, warn_g = drive(got)
    assert err_g == err_r and len(warn_g) == len(warn_r)
    if policy == "error":
        assert err_g == "All particle weights are zero."
    else:
        assert len(got.normalization_record) == len(ref.normalization_record)
        w_r, w_g = ref.particle_weights, got.particle_weights
        assert np.array_equal(np.isnan(w_g), np.isnan(w_r))
        np.testing.assert_allclose(w_g[~np.isnan(w_g)], w_r[~np.isnan(w_r)], rtol=1e-12)


def test_properties_settle_pending_work():
    inp, ref, got = _pair(300, 5, True, 8)
    for k in range(5):
        got.update(int(inp['outcomes'][k]), np.array([inp['ts'][k]]))
    assert len(got._queue) == 5 and got._pending is None                  # buffered, nothing launched yet
    assert len(got.normalization_record) == 5                             # reading a record flushes
    assert not got._queue and got._pending is None
    assert got.data_record == [int(o) for o in inp['outcomes'][:5]]


def test_risk_and_information_gain_host_algebra():
    """SMCUpdater.bayes_risk / expected_information_gain (smc.py:553-657): the host recombines the per-outcome
    reductions (normalisation, first and second moment about the mean, KL sums) into the reference's numbers."""
    inp, ref, got = _pair(700, 9, False, None, resample_thresh=0.0)
    for k in range(6):
        for u in (ref, got):
            u.update(int(inp['outcomes'][k]), np.array([inp['ts'][k]]))
    ts = np.array([0.4, 2.0, 9.0, 33.0])
    np.testing.assert_allclose(got.bayes_risk(ts), ref.bayes_risk(ts), rtol=1e-10)
    np.testing.assert_allclose(got.expected_information_gain(ts), ref.expected_information_gain(ts), rtol=1e-10)
    assert got.risk(2.0).shape == (1,)
