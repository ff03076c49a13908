"""Scratch micro-benchmark of the binned resample's stages (not part of the judged bench):
    python tools/bench_resample.py [n] [d]
CUDA-event times of pass 1 (sums + moments), pass 2 (counts), pass 3 (move) back to back, and of a whole
SMCUpdater.resample() (host work included) for the binned and the guided draw."""
import os
import sys
import warnings
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-qinfer_b200"))
import qinfer_b200 as qb
from qinfer_b200 import _lib
from qinfer_b200.engine import _ptr, _stream


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


class Fixed(object):
    def __init__(self, x):
        self.x, self.n_rvs = x, x.shape[1]

    def sample(self, n=1):
        return self.x


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 7
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rs = np.random.RandomState(0)
    if d == 1:
        model, x = qb.SimplePrecessionModel(), rs.random_sample((n, 1)) * 0.2 + 0.4
    else:
        model = qb.RandomizedBenchmarkingModel()
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.3 * rs.random_sample(n), 0.3 * rs.random_sample(n)])
    w = rs.random_sample(n) ** 2
    w /= w.sum()
    for draw in ("binned", "guided"):
        res = qb.LiuWestResampler(a=0.98, rng='philox', seed=3, scan='fast', draw=draw)
        up = qb.SMCUpdater(model, n, Fixed(x), resampler=res)
        up.particle_weights = w
        cloud = up._cloud
        if draw == "binned":
            cloud.preallocate_binned()
            lib = cloud.lib
            mean = np.dot(w, x)
            S = 0.01 * np.eye(d)
            wsb = cloud._bin_ws.numel() * 8
            t1 = timed(lambda: lib.qb_lw_binned_sums(_ptr(cloud.x), _ptr(cloud.w), _ptr(cloud.stats), n, d,
                                                     _ptr(cloud.moments_out), None, 0.0, _ptr(cloud._bin_ws), wsb, _stream()))
            # (pass 2 accumulates into counts that pass 1 zeroes: always time them as a pair)
            t12 = timed(lambda: cloud.binned_prepare(n, 3, 0))
            t2 = t12 - t1
            t3 = timed(lambda: cloud.binned_move(mean, S, 0.98, 3, 0, 5, 0, n, False, fuse_weights=True))
            by = 8.0 * n
            print("n=%d d=%d binned: sums+moments %.1f us (%.0f GB/s)  counts %.1f us  move %.1f us (%.0f GB/s)" % (
                n, d, t1, by * (d + 1) / t1 / 1e3, t2, t3, by * (2 * d + 2) / t3 / 1e3))

        import time
        trace = []
        if draw == "binned" and os.environ.get("QB_TRACE"):
            for name in ("binned_resample", "binned_prepare", "binned_moments_wait", "binned_move", "binned_counters_wait", "binned_retry_wait", "adopt_binned"):
                def wrap(fn, name=name):
                    def inner(*a, **k):
                        t0 = time.perf_counter()
                        r = fn(*a, **k)
                        trace.append((name, t0, time.perf_counter()))
                        return r
                    return inner
                setattr(cloud, name, wrap(getattr(cloud, name)))
        ts = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for rep in range(8):
                up.particle_weights = w
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tstart = time.perf_counter()
                del trace[:]
                up.resample()
                tend = time.perf_counter()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
                if trace and rep == 7:
                    print("   host trace (us from start): " + "  ".join("%s %.0f-%.0f" % (nm, (a - tstart) * 1e6, (b - tstart) * 1e6)
                                                                  for nm, a, b in trace) + "  end %.0f" % ((tend - tstart) * 1e6))
        print("n=%d d=%d draw=%s whole resample(): median %.1f us  (all: %s)" % (
            n, d, draw, float(np.median(ts[2:])), " ".join("%.0f" % t for t in ts)))


if __name__ == "__main__":
    main()
