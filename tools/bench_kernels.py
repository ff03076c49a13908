"""Scratch micro-benchmarks of individual kernels (not part of the judged bench):
    python tools/bench_kernels.py update [n] | resample [n] | all
Times back-to-back launches with CUDA events and prints achieved algorithmic GB/s."""
import os
import sys
import ctypes
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-qinfer_b200"))
import qinfer_b200 as qb
from qinfer_b200 import _lib
from qinfer_b200.engine import DeviceCloud, _ptr, _stream


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def bench_update(n, model, ep_setup, d, label, fuse_list=(1,), fast_math=False):
    desc = qb.describe_model(model)
    desc.c_model.fast_math = 1 if fast_math else 0
    cloud = DeviceCloud(desc, n)
    rs = np.random.RandomState(0)
    if d == 1:
        x = rs.random_sample((n, 1))
    elif d == 3:
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n), 0.5 * rs.random_sample(n), 0.5 * rs.random_sample(n)])
    else:
        x = rs.random_sample((n, d)) / d
    cloud.upload_locations(x)
    cloud.set_uniform_weights()
    print("  x %#x  w0 %#x  w1 %#x" % (cloud.x.data_ptr(), cloud._w[0].data_ptr(), cloud._w[1].data_ptr()))
    ep = _lib.QbExpparams()
    outcome = ep_setup(ep)
    state = {"src": 0}

    for k in fuse_list:
        steps = [(ep, outcome, False)] * k

        def step():
            cloud.fused_update(steps, state["src"])
            state["src"] ^= 1
        us = timed(step, reps=100)
        gb = 8.0 * (d + 2) * n / (us * 1e-6) / 1e9
        print("%-28s n=%d K=%d %8.1f us/launch %7.1f us/update %7.1f GB/s per launch (%.1f%% of 6540)  %.3g pu/s"
              % (label, n, k, us, us / k, gb, gb / 65.40, n * k / (us * 1e-6)))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 7
    if what == "update1":
        def prec1(ep):
            ep.t = 17.3
            return 1
        print("QB_UPD_ZIGZAG=%s QB_UPD_L2HINT=%s QB_ALLOC_PAD=%s" % (os.environ.get("QB_UPD_ZIGZAG"),
              os.environ.get("QB_UPD_L2HINT"), os.environ.get("QB_ALLOC_PAD")))
        bench_update(n, qb.SimplePrecessionModel(), prec1, 1, "update precession d=1", (1, 8))
        return
    if what == "resample1":
        model = qb.SimplePrecessionModel()
        cloud = DeviceCloud(qb.describe_model(model), n)
        rs = np.random.RandomState(0)
        cloud.upload_locations(rs.random_sample((n, 1)) * 0.2 + 0.4)
        w = rs.random_sample(n) ** 4
        cloud.upload_weights(w / w.sum())
        mean, S = np.full(1, 0.5), np.eye(1) * 0.01
        print("--- device-RNG resample kernels d=1 n=%d" % n)
        print("  moments                  %8.1f us" % timed(lambda: cloud.lib.qb_moments(_ptr(cloud.x), _ptr(cloud.w), _ptr(cloud.stats), n, 1, _ptr(cloud.moments_out), _ptr(cloud.ws), cloud.ws_bytes, _stream()), 10))
        print("  cdf fast                 %8.1f us" % timed(lambda: cloud.cdf(_lib.QB_SCAN_FAST), 10))
        print("  cdf fast+guide           %8.1f us" % timed(lambda: cloud.cdf(_lib.QB_SCAN_FAST_GUIDE), 10))
        print("  fused draw+move (guided) %8.1f us" % timed(lambda: cloud.lw_draw_move(mean, S, 0.98, 1, 0, 2, 0, n, True), 10))
        print("  merge draw+move (sorted) %8.1f us  (3 launches)" % timed(lambda: cloud.lw_merge_move(mean, S, 0.98, 1, 0, 2, 0, n, True), 10))
        return
    if what in ("update", "all"):
        def prec(ep):
            ep.t = 17.3
            return 1
        bench_update(n, qb.SimplePrecessionModel(), prec, 1, "update precession d=1", (1, 2, 4, 8))

        def rbb(ep):
            ep.m = 37
            ep.n_meas = 25
            return 12
        bench_update(n // 4, qb.BinomialModel(qb.RandomizedBenchmarkingModel()), rbb, 3, "update binomial(RB) d=3", (1, 8))
        bench_update(n // 4, qb.BinomialModel(qb.RandomizedBenchmarkingModel()), rbb, 3,
                     "update binomial(RB) fast_math", (1, 8), fast_math=True)

        def rb(ep):
            ep.m = 37
            return 1
        bench_update(n // 4, qb.RandomizedBenchmarkingModel(), rb, 3, "update RB d=3")
        bench_update(n // 4, qb.RandomizedBenchmarkingModel(), rb, 3, "update RB d=3 fast_math", fast_math=True)

        def tomo(ep):
            for c in range(16):
                ep.meas[c] = 0.1 * (c % 3)
            ep.meas[0] = 1.0
            return 1
        bench_update(n // 8, qb.TomographyModel(qb.pauli_basis(2)), tomo, 16, "update tomography d=16")
    if what in ("resample", "all"):
        for d, model in ((1, qb.SimplePrecessionModel()), (3, qb.RandomizedBenchmarkingModel()),
                         (16, qb.TomographyModel(qb.pauli_basis(2)))):
            nn = n if d == 1 else n // 8
            desc = qb.describe_model(model)
            cloud = DeviceCloud(desc, nn)
            rs = np.random.RandomState(0)
            x = rs.random_sample((nn, d)) * 0.2 + 0.4
            cloud.upload_locations(x)
            w = rs.random_sample(nn) ** 4
            cloud.upload_weights(w / w.sum())
            cloud.preallocate_resample()
            print("--- resample kernels d=%d n=%d" % (d, nn))
            print("  moments      %8.1f us" % timed(lambda: cloud.lib.qb_moments(_ptr(cloud.x), _ptr(cloud.w), _ptr(cloud.stats), nn, d, _ptr(cloud.moments_out), _ptr(cloud.ws), cloud.ws_bytes, _stream()), 20))
            print("  cdf fast     %8.1f us" % timed(lambda: cloud.cdf(_lib.QB_SCAN_FAST), 20))
            print("  cdf exact    %8.1f us  (fell back: %d)" % (timed(lambda: cloud.cdf(_lib.QB_SCAN_EXACT), 10, 2),
                                                            cloud.exact_scan_fell_back()))
            cloud.rng_uniform(cloud._u, nn, 1, 0)
            print("  rng uniform  %8.1f us" % timed(lambda: cloud.rng_uniform(cloud._u, nn, 1, 0), 20))
            print("  rng normal   %8.1f us" % timed(lambda: cloud.rng_normal(cloud._eps, nn * d, 1, 0), 20))
            if d == 1:
                print("  mt19937 uniform %8.1f us" % timed(lambda: cloud.mt19937_uniform(cloud._u, nn), 3))
                print("  mt19937 normal  %8.1f us" % timed(lambda: cloud.mt19937_normal(cloud._eps, nn * d), 3))
            print("  draw         %8.1f us" % timed(lambda: cloud.draw(cloud._u, nn), 20))
            mean = np.full(d, 0.5)
            S = np.eye(d) * 0.01
            print("  lw_move      %8.1f us" % timed(lambda: cloud.lw_move(mean, S, 0.98, cloud._eps, nn, True), 20))
            if d == 16:
                print("  canonicalize %8.1f us" % timed(lambda: cloud.canonicalize(), 10))
            if d <= 4:
                print("  cdf fast+guide %6.1f us" % timed(lambda: cloud.cdf(_lib.QB_SCAN_FAST_GUIDE), 20))
                print("  fused draw+move (guided) %8.1f us" % timed(
                    lambda: cloud.lw_draw_move(mean, S, 0.98, 1, 0, 2, 0, nn, True), 20))
                print("  merge draw+move (sorted) %8.1f us  (3 launches)" % timed(
                    lambda: cloud.lw_merge_move(mean, S, 0.98, 1, 0, 2, 0, nn, True), 20))


if __name__ == "__main__":
    main()
