import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, "python-qinfer_b200")
import bench_kernels as bk, qinfer_b200 as qb
def prec(ep):
    ep.t = 17.3
    return 1
bk.bench_update(10**7, qb.SimplePrecessionModel(), prec, 1, "fused", (8,))
