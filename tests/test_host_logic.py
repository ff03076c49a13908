"""Host-side logic that needs no GPU: model recognition, expparams marshalling, loud failure."""
import numpy as np
import pytest
import torch

import qinfer_b200 as qb
from qinfer_b200 import _lib


def test_describe_builtin_models():
    d = qb.describe_model(qb.SimplePrecessionModel(min_freq=0.25))
    assert (d.kind, d.d, d.binomial, d.min_freq) == (_lib.QB_MODEL_PRECESSION, 1, False, 0.25)
    d = qb.describe_model(qb.RandomizedBenchmarkingModel(interleaved=True))
    assert (d.kind, d.d, d.interleaved) == (_lib.QB_MODEL_RB, 4, True)
    d = qb.describe_model(qb.BinomialModel(qb.RandomizedBenchmarkingModel()))
    assert (d.kind, d.d, d.binomial, d.binomial_scalar) == (_lib.QB_MODEL_RB, 3, True, False)
    d = qb.describe_model(qb.BinomialModel(qb.SimplePrecessionModel()))
    assert d.binomial and d.binomial_scalar
    d = qb.describe_model(qb.TomographyModel(qb.pauli_basis(2)))
    assert (d.kind, d.d, d.dim) == (_lib.QB_MODEL_TOMOGRAPHY, 16, 4)
    assert d.basis.shape == (16, 4, 4)
    d = qb.describe_model(qb.MLEModel(qb.BinomialModel(qb.RandomizedBenchmarkingModel()), 2.5))
    assert (d.kind, d.binomial, d.likelihood_power, d.c_model.likelihood_power) == (_lib.QB_MODEL_RB, True, 2.5, 2.5)
    assert qb.describe_model(qb.SimplePrecessionModel()).c_model.likelihood_power == 1.0


def test_recognition_works_on_foreign_objects_with_the_reference_layout():
    """A real qinfer model is recognised by class name + the attributes SURVEY §8b lists."""
    import smc_oracle as o     # stands in for `qinfer` (same class names and private attributes)
    d = qb.describe_model(o.BinomialModel(o.RandomizedBenchmarkingModel()))
    assert (d.kind, d.binomial) == (_lib.QB_MODEL_RB, True)
    d = qb.describe_model(o.TomographyModel(o.pauli_basis(1)))
    assert (d.kind, d.d) == (_lib.QB_MODEL_TOMOGRAPHY, 4)


def test_unsupported_model_raises_no_cpu_fallback():
    class PoissonModel(object):
        n_modelparams = 1
    with pytest.raises(qb.UnsupportedModelError, match="no CPU fallback"):
        qb.describe_model(PoissonModel())
    assert qb.describe_model(qb.CoinModel()).kind == 4          # QB_MODEL_COIN: the reference tests' risk / gain model
    assert qb.describe_model(qb.BinomialModel(qb.CoinModel())).binomial

    class MyPrecession(qb.SimplePrecessionModel):      # subclasses may override likelihood: not silently accepted
        pass
    with pytest.raises(qb.UnsupportedModelError):
        qb.describe_model(MyPrecession())


def test_expparams_marshalling():
    m = qb.BinomialModel(qb.RandomizedBenchmarkingModel())
    ep = np.empty((2,), dtype=m.expparams_dtype)
    ep['m'] = [3, 800]
    ep['n_meas'] = [25, 10]
    r = qb.describe_model(m).expparams_record(ep, 1)
    assert (r.m, r.n_meas) == (800, 10)
    m = qb.BinomialModel(qb.SimplePrecessionModel())
    ep = np.empty((1,), dtype=m.expparams_dtype)
    ep['x'] = 2.5
    ep['n_meas'] = 7
    r = qb.describe_model(m).expparams_record(ep, 0)
    assert (r.t, r.w_, r.n_meas) == (2.5, 0.0, 7)
    r = qb.describe_model(qb.SimplePrecessionModel()).expparams_record(np.array([1.25]), 0)
    assert r.t == 1.25
    inv = qb.SimpleInversionModel()
    ep = np.empty((1,), dtype=inv.expparams_dtype)
    ep['t'], ep['w_'] = 3.0, 0.5
    r = qb.describe_model(inv).expparams_record(ep, 0)
    assert (r.t, r.w_) == (3.0, 0.5)
    tm = qb.TomographyModel(qb.pauli_basis(1))
    ep = np.empty((1,), dtype=tm.expparams_dtype)
    ep['meas'][0] = [0.5, 0.1, 0.2, 0.3]
    r = qb.describe_model(tm).expparams_record(ep, 0)
    assert list(r.meas[:4]) == [0.5, 0.1, 0.2, 0.3]
    with pytest.raises(ValueError):
        qb.describe_model(qb.BinomialModel(qb.RandomizedBenchmarkingModel())).expparams_record(np.array([1.0]), 0)


def test_bases_match_oracle():
    import smc_oracle as o
    for nq in (1, 2):
        assert np.array_equal(qb.pauli_basis(nq).data, o.pauli_basis_data(nq))
    assert np.array_equal(qb.gell_mann_basis(3).data, o.gell_mann_basis_data(3))


def test_liu_west_parameters():
    r = qb.LiuWestResampler(a=0.9)
    assert r.h == np.sqrt(1 - 0.81)
    r.a = 0.5
    assert r.h == np.sqrt(0.75)
    r = qb.LiuWestResampler(a=0.9, h=0.1)
    r.a = 0.5
    assert r.h == 0.1                               # explicit h survives `a` updates (resamplers.py:248-252)
    with pytest.raises(ValueError):
        qb.LiuWestResampler(rng='philox', kernel=lambda *s: np.zeros(s))


def test_sqrtm_psd_matches_reference_test():
    """tests/test_utils.py:132-152 restated: sqrtm_psd(Y)^2 == Y, also for singular Y."""
    rs = np.random.RandomState(0)
    X = rs.randn(5, 5)
    Y = X @ X.T
    S, err = qb.sqrtm_psd(Y)
    np.testing.assert_allclose(S @ S, Y, atol=1e-10)
    assert err < 1e-10
    Y[:, 0] = 0
    Y[0, :] = 0
    S = qb.sqrtm_psd(Y, est_error=False)
    np.testing.assert_allclose(S @ S, Y, atol=1e-10)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_path_fails_loudly_without_cuda():
    with pytest.raises(_lib.QbError, match="no CPU fallback"):
        qb.SMCUpdater(qb.SimplePrecessionModel(), 100, qb.UniformDistribution([0, 1]))
    with pytest.raises(_lib.QbError, match="no CPU fallback"):
        qb.SimplePrecessionModel().likelihood(np.array([0]), np.array([[0.5]]), np.array([1.0]))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; the package must not reference it."""
    import os
    pkg = os.path.dirname(qb.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "smc_oracle" not in src and "import oracle" not in src, fn


def test_sqrtm_psd_one_by_one_shortcut_equals_eigh():
    """The 1 x 1 shortcut returns exactly what the scipy.linalg.eigh route of utils.py:593-607 returns."""
    import scipy.linalg
    for a in (0.0, 1e-300, 3.7e-9, 0.25, 2.0, 1e300, -1e-12):
        A = np.array([[a]])
        w, v = scipy.linalg.eigh(A)
        w[w <= 0] = 0
        np.sqrt(w, out=w)
        want = (v * w).dot(v.conj().T)
        want_err = np.linalg.norm(np.dot(want, want) - A, 'fro')
        got, err = qb.sqrtm_psd(A)
        assert got.shape == (1, 1) and got[0, 0] == want[0, 0] and err == want_err
    with pytest.raises(ValueError):
        qb.sqrtm_psd(np.array([[np.nan]]))


def test_nvtx_ranges_are_opt_in():
    """QB_NVTX=1 wraps the host entry points in NVTX ranges; unset, the functions are untouched (no hot-path cost)."""
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import qinfer_b200 as qb; from qinfer_b200.engine import DeviceCloud; "
            "print(int(hasattr(qb.SMCUpdater.update, '__wrapped__')), int(hasattr(qb.SMCUpdater.resample, '__wrapped__')),"
            " int(hasattr(DeviceCloud.fused_update, '__wrapped__')))") % os.path.join(ROOT, "python-qinfer_b200")
    for flag, want in (("1", "1 1 1"), ("0", "0 0 0")):
        env = dict(os.environ, QB_NVTX=flag)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert out.stdout.strip().splitlines()[-1] == want
