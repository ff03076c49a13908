"""Supplementary measurements of the other BASELINE.json configurations (C1, C3, C4) through the public API,
with the NumPy oracle port of the reference timed beside each on a bounded sample.  Not the judged bench
(bench.py measures C2); the output goes to profiles/ as context for DESIGN.md.  It lives under tests/ because it
times the oracle (test infrastructure) beside the device path; pytest does not collect it.

    python tests/bench_configs.py [--quick]
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "python-qinfer_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import qinfer_b200 as qb          # noqa: E402
import smc_oracle as oracle       # noqa: E402
import cases                      # noqa: E402


def timed_run(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def c1(ns, n, n_updates, lazy):
    inp = cases.precession_inputs(n_particles=n, n_updates=n_updates)
    kw = dict(lazy=True) if lazy else {}
    np.random.seed(0)
    up = ns.SMCUpdater(ns.SimplePrecessionModel(), n, cases.FixedPrior(inp['prior']), **kw)

    def run():
        for k in range(n_updates):
            up.update(int(inp['outcomes'][k]), np.array([inp['ts'][k]]))
        return up.est_mean(), up.resample_count
    run.updater = up
    return run


def c3(ns, n, n_updates, lazy, fast_math=False):
    inp = cases.rb_inputs(n_particles=n, n_updates=n_updates)
    model = ns.BinomialModel(ns.RandomizedBenchmarkingModel())
    kw = dict(lazy=True, resampler=ns.LiuWestResampler(a=0.98, rng='philox', scan='fast', seed=1)) if lazy else {}
    if fast_math:
        kw['fast_math'] = True
    np.random.seed(0)
    up = ns.SMCUpdater(model, n, cases.FixedPrior(inp['prior']), **kw)
    eps = np.empty((n_updates,), dtype=model.expparams_dtype)
    eps['m'] = inp['ms']
    eps['n_meas'] = inp['n_meas']

    def run():
        up.batch_update(inp['counts'], eps, resample_interval=1)
        return up.est_mean(), up.resample_count
    run.updater = up
    return run


def c4(ns, n, n_updates, lazy):
    basis = ns.pauli_basis(2)
    inp = cases.tomography_inputs(np.asarray(basis.data), n_particles=n, n_updates=n_updates)
    model = ns.TomographyModel(basis)
    kw = dict(lazy=True, resampler=ns.LiuWestResampler(a=0.98, rng='philox', scan='fast', seed=1)) if lazy else {}
    np.random.seed(0)
    up = ns.SMCUpdater(model, n, cases.FixedPrior(inp['prior']), **kw)
    eps = []
    for k in range(n_updates):
        ep = np.empty((1,), dtype=model.expparams_dtype)
        ep['meas'][0] = inp['meas'][k]
        eps.append(ep)

    def run():
        for k in range(n_updates):
            up.update(int(inp['outcomes'][k]), eps[k])
        return up.est_mean(), up.resample_count
    run.updater = up
    return run


def main():
    quick = "--quick" in sys.argv
    gpu_ns = cases.Namespace(SMCUpdater=qb.SMCUpdater, LiuWestResampler=qb.LiuWestResampler,
                             SimplePrecessionModel=qb.SimplePrecessionModel,
                             RandomizedBenchmarkingModel=qb.RandomizedBenchmarkingModel, BinomialModel=qb.BinomialModel,
                             TomographyModel=qb.TomographyModel, pauli_basis=qb.pauli_basis)
    cpu_ns = cases.oracle_namespace()
    plan = [
        ("C1 SimplePrecession N=1e3 x 100 updates (default LW, numpy RNG, exact scan)", c1, 1000, 100, 1000, 100, False),
        ("C1 as above with lazy=True (updates buffered and fused 8 per launch: the small-cloud path)", c1, 1000, 100, 1000,
         100, True),
        ("C3 Binomial(RB) N=1e6 x 201 updates, batch_update(resample_interval=1)", c3, 10 ** 6, 201, 10 ** 5, 40, True),
        ("C3 as above with fast_math=True (integer powers instead of pow / log / exp)",
         lambda ns, n, k, lazy: c3(ns, n, k, lazy, fast_math=(ns is gpu_ns)), 10 ** 6, 201, 10 ** 5, 40, True),
        ("C4 Tomography 2 qubits (d=16) N=1e6 x 200 updates, canonicalize", c4, 10 ** 6, 200, 2000, 40, True),
    ]
    results = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for label, fn, n, k, n_cpu, k_cpu, lazy in plan:
            if quick:
                n, k = min(n, 10 ** 5), min(k, 40)
            warm = fn(gpu_ns, min(n, 20000), 10, lazy)                 # warm-up: load every kernel of the path,
            warm()                                                     # including the resample's
            warm.updater.resample()
            warm.updater.est_mean()
            fn(gpu_ns, n, min(k, 30), lazy)()                          # ... and the allocator / pinned blocks at full size
            t_gpu, (mean_g, rc_g) = timed_run(fn(gpu_ns, n, k, lazy))
            t_cpu, (mean_c, rc_c) = timed_run(fn(cpu_ns, n_cpu, k_cpu, False))
            r = dict(config=label, gpu_particles=n, gpu_updates=k, gpu_seconds=t_gpu, gpu_resamples=int(rc_g),
                     gpu_particle_updates_per_s=n * k / t_gpu,
                     cpu_particles=n_cpu, cpu_updates=k_cpu, cpu_seconds=t_cpu, cpu_resamples=int(rc_c),
                     cpu_particle_updates_per_s=n_cpu * k_cpu / t_cpu,
                     speedup=(n * k / t_gpu) / (n_cpu * k_cpu / t_cpu))
            results.append(r)
            print(json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
