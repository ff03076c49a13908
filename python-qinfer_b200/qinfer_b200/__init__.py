"""qinfer_b200 — B200-native SMC particle-filter engine behind QInfer's
SMCUpdater / Model / Resampler plugin surface (hot path only)."""
from ._exceptions import ApproximationWarning, ResamplerError, ResamplerWarning, UnsupportedModelError
from .distributions import (GinibreTomographyPrior, MultivariateNormalDistribution, ParticleDistribution,
                            PostselectedDistribution, UniformDistribution)
from .models import (BinomialModel, CoinModel, DiffusiveTomographyModel, GaussianRandomWalkModel, IntegerDomain, MLEModel,
                     Model, PoisonedModel, RandomWalkModel, RandomizedBenchmarkingModel, SimpleInversionModel,
                     SimplePrecessionModel, TomographyBasis, TomographyModel, describe_model, gell_mann_basis,
                     pauli_basis)
from .resamplers import LiuWestResampler, Resampler, sqrtm_psd
from .smc import SMCUpdater
from .simple_est import simple_est_prec, simple_est_rb

__all__ = [
    'ApproximationWarning', 'ResamplerError', 'ResamplerWarning', 'UnsupportedModelError',
    'GinibreTomographyPrior', 'MultivariateNormalDistribution', 'ParticleDistribution', 'PostselectedDistribution', 'UniformDistribution',
    'BinomialModel', 'CoinModel', 'DiffusiveTomographyModel', 'GaussianRandomWalkModel', 'PoisonedModel', 'RandomWalkModel',
    'IntegerDomain', 'MLEModel', 'Model', 'RandomizedBenchmarkingModel', 'SimpleInversionModel', 'SimplePrecessionModel',
    'TomographyBasis', 'TomographyModel', 'describe_model', 'gell_mann_basis', 'pauli_basis',
    'LiuWestResampler', 'Resampler', 'sqrtm_psd', 'SMCUpdater', 'simple_est_prec', 'simple_est_rb',
]
