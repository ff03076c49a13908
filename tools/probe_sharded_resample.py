"""Scratch: where does a sharded resample spend its time? torchrun --nproc-per-node 2 tools/probe_sharded_resample.py"""
import os, sys, time, warnings
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-qinfer_b200"))
import qinfer_b200 as qb
from qinfer_b200 import sharded
from qinfer_b200.sharded import ShardedSMCUpdater

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dist.init_process_group("nccl")
n = 10 ** 7
class P:
    n_rvs = 1
    def sample(self, n=1): return np.random.RandomState(rank).random_sample((n, 1))
T = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n * world, P(), resampler=qb.LiuWestResampler(rng='philox', scan='fast', seed=1), lazy=True)
    for k in range(8): up.update(k % 2, np.array([1.5 ** k]))
    up.resample(); up.resample()
    for nm in ("classify", "bucket", "local_draw", "gather_rows"): wrap(up._ops, nm)
    for nm in ("all_gather_scalars", "exchange_counts", "all_to_all_v", "all_reduce_sum"): wrap(up._comm, nm)
    for nm in ("cdf", "lw_move", "rng_uniform", "rng_normal", "read_counter", "set_uniform_weights"): wrap(up._cloud, nm)
    wrap(up, "_global_moments")
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(5): up.resample()
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 5
    if rank == 0:
        print("total per resample %.3f ms (with per-phase syncs)" % (tot * 1e3))
        for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print("  %-22s %.3f ms" % (k, v / 5 * 1e3))
    up.close()
dist.destroy_process_group()
