"""One-call estimators on the B200 engine — the callers either side of the hot path that SURVEY §8(f1) lists
(``qinfer.simple_est``: ``simple_est_prec``, ``simple_est_rb``, ``do_update``, ``data_to_params``,
``load_data_or_txt``; simple_est.py:69-260).  Same arguments, column conventions and return values; the posterior is
computed by ``SMCUpdater.batch_update(resample_interval=1)``, i.e. by fused update launches.
"""
import numpy as np

from .distributions import PostselectedDistribution, UniformDistribution
from .models import BinomialModel, RandomizedBenchmarkingModel, SimplePrecessionModel
from .smc import SMCUpdater

try:  # pandas is optional, exactly as in the reference
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None


def data_to_params(data, expparams_dtype, col_outcomes=(0, 'counts'), cols_expparams=None):
    """Split a data table into the outcome column and an ``expparams`` array (simple_est.py:69-104).
    Columns are given as ``(index, name)`` pairs: the index is used for plain 2-D arrays, the name for record arrays."""
    by_index = data.dtype.fields is None

    def col(spec):
        return data[..., spec[0]] if by_index else data[spec[1]]

    outcomes = col(col_outcomes).astype(int)
    expparams = np.empty(outcomes.shape, dtype=expparams_dtype)
    if isinstance(expparams_dtype, str) or np.dtype(expparams_dtype).fields is None:
        expparams[:] = col(cols_expparams)
    else:
        for key, spec in cols_expparams.items():
            expparams[key] = col(spec)
    return outcomes, expparams


def load_data_or_txt(data, dtype):
    """An array, a pandas DataFrame, or a CSV file name / file object (simple_est.py:106-117)."""
    if isinstance(data, np.ndarray):
        return data
    if pd is not None and isinstance(data, pd.DataFrame):
        return data.to_records(index=False)
    if hasattr(data, 'read') or isinstance(data, str):
        return np.loadtxt(data, dtype=dtype, delimiter=',')
    raise TypeError("Expected a filename, an array or a file-like object.")


def do_update(model, n_particles, prior, outcomes, expparams, return_all, resampler=None):
    """simple_est.py:120-139."""
    updater = SMCUpdater(model, n_particles, prior, resampler=resampler)
    updater.batch_update(outcomes, expparams, resample_interval=1)
    mean = updater.est_mean()
    cov = updater.est_covariance_mtx()
    if model.n_modelparams == 1:
        mean = mean[0]
        cov = cov[0, 0]
    if not return_all:
        return mean, cov
    return mean, cov, {'updater': updater}


def simple_est_prec(data, freq_min=0.0, freq_max=1.0, n_particles=6000, return_all=False):
    """Precession-frequency estimate from rows of (counts, t, n_shots) (simple_est.py:141-189)."""
    model = BinomialModel(SimplePrecessionModel(freq_min))
    prior = UniformDistribution([0, freq_max])
    data = load_data_or_txt(data, [('counts', 'uint'), ('t', float), ('n_shots', 'uint')])
    outcomes, expparams = data_to_params(data, model.expparams_dtype,
                                         cols_expparams={'x': (1, 't'), 'n_meas': (2, 'n_shots')})
    return do_update(model, n_particles, prior, outcomes, expparams, return_all)


def simple_est_rb(data, interleaved=False, p_min=0.0, p_max=1.0, n_particles=8000, return_all=False):
    """Randomized-benchmarking estimate from rows of (counts, m, n_shots[, reference]) (simple_est.py:191-260)."""
    model = BinomialModel(RandomizedBenchmarkingModel(interleaved=interleaved))
    box = [[p_min, p_max], [0, 1], [0, 1]] if not interleaved else [[p_min, p_max], [p_min, p_max], [0, 1], [0, 1]]
    prior = PostselectedDistribution(UniformDistribution(box), model)
    data = load_data_or_txt(data, [('counts', 'uint'), ('m', 'uint'), ('n_shots', 'uint')] +
                            ([('reference', 'uint')] if interleaved else []))
    cols = {'m': (1, 'm'), 'n_meas': (2, 'n_shots')}
    if interleaved:
        cols['reference'] = (3, 'reference')
    outcomes, expparams = data_to_params(data, model.expparams_dtype, cols_expparams=cols)
    return do_update(model, n_particles, prior, outcomes, expparams, return_all)
