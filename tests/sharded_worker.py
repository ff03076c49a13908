"""Worker for tests/test_gpu_sharded.py: run under torchrun with 2+ GPUs (NCCL)."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "python-qinfer_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import qinfer_b200 as qb                                   # noqa: E402
from qinfer_b200.sharded import ShardedSMCUpdater, ShardLayout   # noqa: E402


class Fixed(object):
    def __init__(self, s):
        self._s = s
        self.n_rvs = s.shape[1]

    def sample(self, n=1):
        assert n == self._s.shape[0]
        return self._s.copy()


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    n_global = 200003
    layout = ShardLayout(n_global, world)
    lo, hi = layout.offsets[rank], layout.offsets[rank + 1]
    rs = np.random.RandomState(5)
    x = rs.random_sample((n_global, 1))
    ts = (9.0 / 8.0) ** np.arange(40)
    outcomes = (rs.random_sample(40) >= np.cos(ts * 0.5 / 2) ** 2).astype(int)
    fails = []

    def check(cond, msg):
        if not cond:
            fails.append(msg)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import smc_oracle as oracle
        np.random.seed(0)
        ou = oracle.SMCUpdater(oracle.SimplePrecessionModel(), n_global, Fixed(x), resample_thresh=0.0)
        for k in range(6):
            ou.update(int(outcomes[k]), np.array([ts[k]]))
        ou.resample()
        om1, oc1 = ou.est_mean(), ou.est_covariance_mtx()
        del ou
        for lazy in (False, True):
            res = qb.LiuWestResampler(rng='philox', scan='fast', seed=11)
            up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n_global, Fixed(x[lo:hi]), resampler=res, lazy=lazy,
                                   resample_thresh=0.0)
            ref = None
            if rank == 0:      # single-GPU engine on the whole cloud as the comparison
                ref = qb.SMCUpdater(qb.SimplePrecessionModel(), n_global, Fixed(x), resample_thresh=0.0)
            # (a) updates: global stats and slab weights equal the single-GPU run
            for k in range(6):
                up.update(int(outcomes[k]), ts[k:k + 1])
                if ref is not None:
                    ref.update(int(outcomes[k]), ts[k:k + 1])
            check(up.n_particles == n_global and up.n_local == hi - lo, "counts")
            if ref is not None:
                check(abs(up.n_ess - ref.n_ess) <= 1e-11 * ref.n_ess, "n_ess %r vs %r" % (up.n_ess, ref.n_ess))
                check(np.allclose(up.normalization_record, ref.normalization_record, rtol=1e-12, atol=0),
                      "normalization record")
                check(np.allclose(up.particle_weights, ref.particle_weights[lo:hi], rtol=1e-11,
                                  atol=1e-15 * ref.particle_weights.max()), "slab weights")
                check(np.allclose(up.est_mean(), ref.est_mean(), rtol=1e-11), "mean")
                check(np.allclose(up.est_covariance_mtx(), ref.est_covariance_mtx(), rtol=1e-8), "cov")
            else:
                up.particle_weights
                up.est_mean()
                up.est_covariance_mtx()
            # hypothetical_update: this slab of the globally normalised hypothetical weights; sample: the same draws
            # everywhere, distributed like the weights
            hyp, hn = up.hypothetical_update(np.array([0, 1]), ts[7:9], return_normalization=True)
            check(hyp.shape == (2, 2, up.n_local) and hn.shape == (2, 2, 1), "hypothetical shapes")
            if ref is not None:
                rh, rn = ref.hypothetical_update(np.array([0, 1]), ts[7:9], return_normalization=True)
                check(np.allclose(hn, rn, rtol=1e-11), "hypothetical normalisation")
                check(np.allclose(hyp, rh[:, :, lo:hi], rtol=1e-10, atol=1e-300), "hypothetical weights")
            smp = up.sample(20000)
            check(smp.shape == (20000, 1), "sample shape %r" % (smp.shape,))
            sm_t = torch.tensor([float(smp.sum())], dtype=torch.float64, device='cuda')
            g2 = [torch.empty_like(sm_t) for _ in range(world)]
            dist.all_gather(g2, sm_t)
            check(all(torch.equal(g2[0], q) for q in g2), "ranks drew different samples")
            mu, sd = up.est_mean()[0], np.sqrt(up.est_covariance_mtx()[0, 0])
            check(abs(smp.mean() - mu) < 6 * sd / np.sqrt(20000), "sample mean %r vs %r" % (smp.mean(), mu))
            # all ranks hold identical global numbers
            t = torch.tensor([up.n_ess, up.normalization_record[-1]], dtype=torch.float64, device='cuda')
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            check(all(torch.equal(g[0], q) for q in g), "ranks disagree on global stats")
            # (b) a forced resample preserves the first two moments and leaves uniform global weights
            m0, c0 = up.est_mean(), up.est_covariance_mtx()
            ess0 = up.n_ess
            up.resample()
            m1, c1 = up.est_mean(), up.est_covariance_mtx()
            check(abs(up.n_ess - n_global) < 1e-6 * n_global, "n_ess after resample %r" % up.n_ess)
            # (offspring below 0 are invalid; the reference's retry re-centres them on an independent draw
            # (resamplers.py:372), which moves ~3 % of the mass away from 0: the oracle's resample is the yardstick)
            check(abs(m1[0] - om1[0]) < 8 * np.sqrt(c0[0, 0] / ess0), "mean moved %r -> %r (oracle %r)" % (m0, m1, om1))
            check(abs(c1[0, 0] / oc1[0, 0] - 1) < 0.05, "cov moved %r -> %r (oracle %r)" % (c0, c1, oc1))
            w = up.particle_weights
            check(np.all(w == 1.0 / n_global), "weights not uniform")
            locs = up.particle_locations
            check(locs.shape == (up.n_local, 1) and w.shape == (up.n_local,) and np.all(locs > 0), "invalid locations")
            # floating slabs: every rank keeps the offspring it drew; the sizes sum to the global count and stay
            # within the slack around the balanced split
            cnt = torch.tensor([up.n_local], dtype=torch.int64, device='cuda')
            dist.all_reduce(cnt)
            check(int(cnt.item()) == n_global, "slab sizes sum to %d" % int(cnt.item()))
            check(abs(up.n_local - (hi - lo)) <= max(4096, 0.03 * (hi - lo)), "slab drifted to %d" % up.n_local)
            check(up.last_exchange == (0, 0), "rows travelled %r" % (up.last_exchange,))
            # ... and the updater keeps working on the resized slab
            up.update(int(outcomes[6]), ts[6:7])
            check(np.isfinite(up.n_ess) and up.n_ess <= n_global, "update after a floated resample")
            up.close()
        # (c) a free-running sharded trajectory lands on the same posterior as the oracle (statistically)
        res = qb.LiuWestResampler(rng='philox', scan='fast', seed=3)
        up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n_global, Fixed(x[lo:hi]), resampler=res, lazy=True)
        for k in range(40):
            up.update(int(outcomes[k]), ts[k:k + 1])
        mean, cov, nres = up.est_mean(), up.est_covariance_mtx(), up.resample_count
        up.close()
        if rank == 0:
            np.random.seed(0)
            ou = oracle.SMCUpdater(oracle.SimplePrecessionModel(), 20000, Fixed(x[:20000]))
            for k in range(40):
                ou.update(int(outcomes[k]), np.array([ts[k]]))
            om, oc = ou.est_mean(), ou.est_covariance_mtx()
            check(nres >= 3, "too few resamples %d" % nres)
            check(abs(mean[0] - om[0]) < 6 * np.sqrt(oc[0, 0]), "posterior mean %r vs oracle %r" % (mean, om))
            check(0.3 < cov[0, 0] / oc[0, 0] < 3.0, "posterior cov %r vs oracle %r" % (cov, oc))
        # (d) PARITY MODE: legacy MT19937 stream continued on the device + the exact scan chained across
        # the slabs -> the global resample indices equal the single-GPU engine's and the reference algorithm's
        # (np.cumsum(w).searchsorted(np.random.random(n), 'right')) bit for bit, retry quirk included
        rs = np.random.RandomState(17)
        problems = []
        n1 = 200003                                                         # slabs >= 32768: the parallel replay scan
        problems.append((qb.SimplePrecessionModel(min_freq=0.35), 0.3 + 0.4 * rs.random_sample((n1, 1)), {}))
        n3 = 50001                                                          # small slabs: the one-lane kernel
        problems.append((qb.RandomizedBenchmarkingModel(),
                         np.column_stack([0.9 + 0.1 * rs.random_sample(n3), 0.6 * rs.random_sample(n3),
                                          0.4 * rs.random_sample(n3)]), {}))
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import cases                                                        # noqa: E402
        basis2 = qb.pauli_basis(2)                                          # d = 16: the generic-d staged kernels
        problems.append((qb.TomographyModel(basis2), cases.ginibre_coords(rs, 20001, np.asarray(basis2.data)),
                         dict(canonicalize=False)))
        for model, xs, kw in problems:
            n_par = xs.shape[0]
            lay = ShardLayout(n_par, world)
            plo, phi = lay.offsets[rank], lay.offsets[rank + 1]
            w = rs.random_sample(n_par) ** 4
            w[rs.randint(0, n_par, n_par // 7)] = 0.0
            w /= w.sum()
            np.random.seed(4242)                                            # the same legacy state on every rank
            up = ShardedSMCUpdater(model, n_par, Fixed(xs[plo:phi]),
                                   resampler=qb.LiuWestResampler(a=0.9, rng='mt19937', scan='exact'), **kw)
            up.particle_weights = w[plo:phi]
            up.resample()
            js = up.last_parity_js.cpu().numpy()
            locs, iters = up.particle_locations, up.resampler.last_n_iters
            after = np.random.random()
            check(np.all(up.particle_weights == 1.0 / n_par), "parity: weights not uniform")
            want = np.cumsum(w).searchsorted(np.random.RandomState(4242).random_sample(n_par), side='right')
            check(np.array_equal(js, want[plo:phi]), "parity: js differ from np.cumsum/searchsorted in %d slots"
                  % int(np.sum(js != want[plo:phi])))
            check(iters > 2 or xs.shape[1] == 16, "parity: the retry loop did not run (%d iterations)" % iters)
            check(bool(np.all(np.asarray(model.are_models_valid(locs)))), "parity: invalid particles left")
            up.close()
            if rank == 0:
                np.random.seed(4242)
                ref = qb.SMCUpdater(model, n_par, Fixed(xs),
                                    resampler=qb.LiuWestResampler(a=0.9, rng='mt19937', scan='exact'), **kw)
                ref.particle_weights = w
                ref.resample()
                check(np.array_equal(ref._cloud._js.cpu().numpy()[plo:phi], js), "parity: js differ from single GPU")
                check(ref.resampler.last_n_iters == iters, "parity: %d iterations vs %d on one GPU"
                      % (iters, ref.resampler.last_n_iters))
                check(np.random.random() == after, "parity: legacy stream position differs from single GPU")
                # (locations: same js, same variates; the global mean / covariance are reduced in another order.
                # Two-qubit states share x_0 = 1/2: the covariance is singular, and the square root of an
                # eigenvalue that is zero up to rounding (+-1e-17) moves by 1e-9 with the summation order)
                check(np.allclose(locs, ref.particle_locations[plo:phi], rtol=1e-12,
                                  atol=1e-7 if xs.shape[1] == 16 else 1e-14),
                      "parity: locations differ from single GPU by %r"
                      % float(np.max(np.abs(locs - ref.particle_locations[plo:phi]))))
        # ... and free-running: the sharded parity trajectory follows the single-GPU parity trajectory
        np.random.seed(7)
        up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n_global, Fixed(x[lo:hi]),
                               resampler=qb.LiuWestResampler(rng='mt19937', scan='exact'))
        for k in range(25):
            up.update(int(outcomes[k]), ts[k:k + 1])
        mean, cov, nres = up.est_mean(), up.est_covariance_mtx(), up.resample_count
        up.close()
        if rank == 0:
            np.random.seed(7)
            ref = qb.SMCUpdater(qb.SimplePrecessionModel(), n_global, Fixed(x),
                                resampler=qb.LiuWestResampler(rng='mt19937', scan='exact'))
            for k in range(25):
                ref.update(int(outcomes[k]), ts[k:k + 1])
            check(nres == ref.resample_count and nres >= 3, "parity trajectory: %d resamples vs %d"
                  % (nres, ref.resample_count))
            check(abs(mean[0] - ref.est_mean()[0]) <= 1e-9 * abs(mean[0]), "parity trajectory: mean %r vs %r"
                  % (mean, ref.est_mean()))
            check(abs(cov[0, 0] / ref.est_covariance_mtx()[0, 0] - 1) <= 1e-7, "parity trajectory: covariance")
    flag = torch.tensor([len(fails)], dtype=torch.int64, device='cuda')
    dist.all_reduce(flag)
    for f in fails:
        print("[rank %d] FAIL: %s" % (rank, f), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
