"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the hot path at small
sizes — lazy speculative + fused updates with roll-back (flag-polled launches, pinned mirrors), the binned resample incl.
its retry kernel, the guided / merge / staged (exact scan, MT19937) resamples, the f3 read-side estimators and the f4
decorators.  Run under torchrun with 2 ranks it also exercises the sharded cloud (IPC mailboxes, in-kernel all-reduce,
floating slabs).

    compute-sanitizer --tool memcheck  python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "python-qinfer_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import qinfer_b200 as qb          # noqa: E402
import cases                      # noqa: E402

warnings.simplefilter("ignore")
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
rs = np.random.RandomState(3 + rank)
n = int(os.environ.get("QB_SAN_N", "40000"))
ts = (9.0 / 8.0) ** np.arange(40)
outcomes = (np.random.RandomState(1).random_sample(40) >= np.cos(ts * 0.5 / 2) ** 2).astype(int)


def drive(up, k=40):
    for i in range(k):
        up.update(int(outcomes[i]), ts[i:i + 1])
    return up.est_mean(), up.resample_count


if world > 1:
    import torch.distributed as dist
    from qinfer_b200.sharded import ShardedSMCUpdater
    dist.init_process_group("nccl")
    for lazy in (False, True):
        up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n * world, cases.FixedPrior(rs.random_sample((n, 1))),
                               resampler=qb.LiuWestResampler(rng='philox', scan='fast', seed=5), lazy=lazy)
        print(rank, "sharded lazy=%s" % lazy, drive(up))
        up.close()
    # parity mode: exact scan chained across the slabs (parallel replay kernel: slabs >= 32768), shared legacy stream
    np.random.seed(12)
    up = ShardedSMCUpdater(qb.SimplePrecessionModel(min_freq=0.05), n * world,
                           cases.FixedPrior(np.random.RandomState(3 + rank).random_sample((n, 1))),
                           resampler=qb.LiuWestResampler(rng='mt19937', scan='exact'))
    print(rank, "sharded parity mode", drive(up, 25))
    up.close()
    dist.destroy_process_group()
    sys.exit(0)

prior = rs.random_sample((n, 1))
for lazy, fuse in ((False, 1), (True, 1), (True, 8)):
    for draw in ("binned", "guided", "merge"):
        res = qb.LiuWestResampler(a=0.98, rng='philox', seed=7, scan='fast', draw=draw)
        up = qb.SMCUpdater(qb.SimplePrecessionModel(min_freq=0.3), n, cases.FixedPrior(0.3 + 0.4 * prior), resampler=res,
                           lazy=lazy, fuse=fuse)
        print("precession lazy=%s fuse=%d draw=%s" % (lazy, fuse, draw), drive(up))
np.random.seed(0)
for rng in ("numpy", "mt19937"):
    up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, cases.FixedPrior(prior),
                       resampler=qb.LiuWestResampler(rng=rng, scan='exact'))
    print("parity mode rng=%s" % rng, drive(up))
# small clouds: the single-CTA parity resample (d = 1 with retries, d = 3)
np.random.seed(2)
up = qb.SMCUpdater(qb.SimplePrecessionModel(min_freq=0.3), 1000, cases.FixedPrior(0.3 + 0.4 * prior[:1000]),
                   resampler=qb.LiuWestResampler(rng='numpy', scan='exact'))
print("small cloud d=1", drive(up))
sm = qb.RandomizedBenchmarkingModel()
x3 = np.column_stack([0.9 + 0.1 * prior[:3000, 0], 0.6 * prior[3000:6000, 0], 0.4 * prior[6000:9000, 0]])
up = qb.SMCUpdater(sm, 3000, cases.FixedPrior(x3), resampler=qb.LiuWestResampler(rng='numpy', scan='exact'))
ep3 = np.empty((1,), dtype=sm.expparams_dtype)
for k in range(12):
    ep3['m'] = 1 + 7 * k
    up.update(k % 2, ep3)
up.resample()
print("small cloud d=3", up.est_mean(), up.resample_count)
# RB under BinomialModel (d = 3, retries), batch_update
inp = cases.rb_inputs(n_particles=n, n_updates=30)
model = qb.BinomialModel(qb.RandomizedBenchmarkingModel())
eps = np.empty((30,), dtype=model.expparams_dtype)
eps['m'], eps['n_meas'] = inp['ms'], inp['n_meas']
up = qb.SMCUpdater(model, n, cases.FixedPrior(inp['prior']), lazy=True,
                   resampler=qb.LiuWestResampler(rng='philox', scan='fast', seed=2))
up.batch_update(inp['counts'], eps, resample_interval=1)
print("rb-binomial", up.est_mean(), up.resample_count, up.est_entropy(), up.est_credible_region(0.9).shape,
      up.est_meanfn(lambda x: x ** 2))
# tomography (d = 16): generic-d kernels, DMMA moments, canonicalize
basis = qb.pauli_basis(2)
tin = cases.tomography_inputs(np.asarray(basis.data), n_particles=8192, n_updates=30)
tm = qb.TomographyModel(basis)
up = qb.SMCUpdater(tm, 8192, cases.FixedPrior(tin['prior']), resampler=qb.LiuWestResampler(rng='philox', scan='fast'))
for k in range(30):
    ep = np.empty((1,), dtype=tm.expparams_dtype)
    ep['meas'][0] = tin['meas'][k]
    up.update(int(tin['outcomes'][k]), ep)
print("tomography", up.resample_count, up.bayes_risk(ep), up.expected_information_gain(ep))
# f4 decorators
np.random.seed(1)
for m in (qb.GaussianRandomWalkModel(qb.SimplePrecessionModel()),
          qb.GaussianRandomWalkModel(qb.SimplePrecessionModel(), fixed_covariance=np.array([1e-6])),
          qb.PoisonedModel(qb.SimplePrecessionModel(), tol=0.01),
          qb.PoisonedModel(qb.SimplePrecessionModel(), n_samples=100, hedge=0.5)):
    d = m.n_modelparams
    x0 = np.column_stack([prior[:, 0]] + ([1e-3 * rs.random_sample(n)] if d == 2 else []))
    for rng in ("numpy", "philox"):
        up = qb.SMCUpdater(m, n, cases.FixedPrior(x0),
                           resampler=qb.LiuWestResampler(rng=rng, scan='exact' if rng == 'numpy' else 'fast', seed=4))
        print(type(m).__name__, rng, drive(up, 20))
dm = qb.DiffusiveTomographyModel(qb.pauli_basis(1))
b1 = np.asarray(qb.pauli_basis(1).data)
x0 = np.column_stack([cases.ginibre_coords(rs, 4096, b1), 0.02 + 0.05 * rs.random_sample(4096)])
up = qb.SMCUpdater(dm, 4096, cases.FixedPrior(x0))
for k in range(12):
    ep = np.empty((1,), dtype=dm.expparams_dtype)
    ep['meas'][0] = [np.sqrt(2) / 2, 0, 0, 0]
    ep['meas'][0][1 + k % 3] = np.sqrt(2) / 2
    ep['t'] = 0.7
    up.update(k % 2, ep)
print("diffusive tomography", up.est_mean(), up.resample_count)
torch.cuda.synchronize()
print("sanitizer workload done")
