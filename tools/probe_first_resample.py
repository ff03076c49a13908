"""Scratch: why is the first resample of a fresh updater slower than the following ones?  (1 GPU)"""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch

be = bench.CudaBackend(0, 1)
n = 10 ** 7
ts, outcomes = bench.make_data(60)
prior = bench.make_prior(n, 99)
warnings.simplefilter("ignore")
bench.process_warmup(be, prior)
for variant in ("plain", "pretouch", "plain", "dummy_resample", "plain_nolazy"):
    up = be.new_updater(n, prior, seed=1000, lazy=(variant != "plain_nolazy"))
    cloud = up._cloud
    if variant == "pretouch":
        cloud.x_alt.zero_()
        cloud._bin_list.zero_()
        cloud._bin_parents.zero_()
        cloud.w_alt.zero_()
    if variant == "dummy_resample":
        up.resample()
        up.reset()
    torch.cuda.synchronize()
    cloud.resample_events = []
    host = []
    orig = up.resample
    def traced():
        t0 = time.perf_counter(); orig(); host.append((time.perf_counter() - t0) * 1e3)
    up.resample = traced
    bench.drive(up, ts, outcomes, 0, 40)
    up._flush()
    torch.cuda.synchronize()
    ev = [a.elapsed_time(b) for a, b in cloud.resample_events]
    print(variant, "event ms:", [round(v, 3) for v in ev], "host ms:", [round(v, 3) for v in host], flush=True)
    be.close(up)
    del up
be.finish()
