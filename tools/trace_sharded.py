"""Scratch: host-side timeline of the first timed steps of a sharded run (torchrun, 2+ GPUs)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
be = bench.CudaBackend(lr, world)
n = 10 ** 7
steps, warm = 20, 5
ts, outcomes = bench.make_data(steps + warm)
prior = bench.make_prior(n, 99 + rank)
warnings.simplefilter("ignore")
bench.process_warmup(be, prior)
for rep in range(2):
    up = be.new_updater(n, prior, seed=1000 + rank)
    bench.drive(up, ts, outcomes, 0, warm)
    up._flush()
    be.barrier()
    marks = []
    stages = []
    def wrap(obj, name):
        fn = getattr(obj, name)
        def inner(*a, **k):
            t_a = time.perf_counter(); r = fn(*a, **k); stages.append((name, t_a, time.perf_counter())); return r
        setattr(obj, name, inner)
    for nm in ("binned_sums", "binned_count", "binned_move", "binned_retry_wait", "binned_counters_wait", "adopt_binned", "_alt_slab"):
        wrap(up._cloud, nm)
    wrap(up._comm, "all_gather_rows")
    wrap(up, "_liu_west_consts")
    orig_resample = up.resample
    def traced():
        t0 = time.perf_counter(); orig_resample(); marks.append(("resample", t0, time.perf_counter()))
    up.resample = traced
    t0 = time.perf_counter()
    stamps = []
    for k in range(warm, warm + steps):
        up.update(int(outcomes[k]), ts[k:k + 1])
        stamps.append(time.perf_counter() - t0)
    up._flush()
    tf = time.perf_counter() - t0
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if rank == 0:
        print("rep", rep, "per-step host stamps (us):", " ".join("%.0f" % (s * 1e6) for s in stamps), "flush %.0f sync %.0f" % (tf * 1e6, te * 1e6))
        print("   resamples:", [(round((a - t0) * 1e6), round((b - t0) * 1e6)) for _, a, b in marks])
        print("   stages:", " ".join("%s %d-%d" % (nm, (a - t0) * 1e6, (b - t0) * 1e6) for nm, a, b in stages))
    be.close(up)
    del up
be.finish()
