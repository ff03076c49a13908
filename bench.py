"""Headline benchmark: particle-updates/s of the SMC Bayes-update + Liu-West path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], "C2"): SimplePrecessionModel, 10^7 particles per
GPU, prior U[0,1], sequential updates with the exp-sparse schedule t_k = (9/8)^(k mod 100),
true omega = 0.5, default LiuWestResampler(a=0.98), resample_thresh = 0.5.  A "step" is
one ``SMCUpdater.update`` call (fused likelihood/weight/normalisation/n_ess kernel, plus the
whole Liu-West resample whenever n_ess < N/2 — 10 to 20 times per 1000 steps).

Metric: particle-updates/s = particles x steps / time, reference convention of
qinfer/perf_testing.py:250-251 (only ``update`` is timed; prior sampling is not).

  value  inputs resident in HBM, timed with CUDA events around the K update calls.
  e2e    the same K updates through the public API starting from HOST arrays: the
         host->device upload of the prior sample and the device->host read-back of the
         posterior (locations, weights, mean) are inside the timed region, as are the
         per-step scalar exchanges (experiment record in, normalisation/n_ess out).
  roofline  fused-update kernel only: algorithmic bytes 8(d+2) per particle
         over its average device duration, measured with CUDA events around each launch.
  cpu_baseline  the NumPy oracle port of the reference, timed on this host on a bounded
         sample of the same workload (first updates of the schedule incl. the resamples).

``--impl reference`` times that CPU port alone for K steps, each step one update over a
bounded particle sample (stated in the output), and prints the same JSON line.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "python-qinfer_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "particle_updates_per_sec"
UNIT = "particle-updates/s"
PARTICLES_PER_GPU = 10 ** 7
TRUE_OMEGA = 0.5


# ---------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------
def schedule(n_steps, start=0):
    k = np.arange(start, start + n_steps)
    return (9.0 / 8.0) ** (k % 100)


def make_data(n_steps, seed=1234, start=0):
    ts = schedule(n_steps, start)
    rs = np.random.RandomState(seed)
    outcomes = (rs.random_sample(n_steps) >= np.cos(ts * TRUE_OMEGA / 2) ** 2).astype(np.int64)
    return ts, outcomes


def make_prior(n, seed):
    return np.random.RandomState(seed).random_sample((n, 1))


class FixedPrior(object):
    n_rvs = 1

    def __init__(self, sample):
        self._s = sample

    def sample(self, n=1):
        return self._s


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/qb_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(n):
    """DRAM bytes (read + write) of ONE fused-update launch from the newest committed `ncu --set full` summary under
    profiles/ (captured at 10^7 particles; scaled linearly to n), or (None, why)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*fused_update_K1_ncu_full.csv")), key=os.path.basename)
    if not files:
        return None, "no ncu --set full summary under profiles/"
    vals = {}
    for row in csv.reader(open(files[-1])):
        if len(row) >= 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(row[1], None)
            if scale is not None:
                vals[row[0]] = float(row[2]) * scale
    if len(vals) != 2:
        return None, "dram metrics missing in %s" % os.path.basename(files[-1])
    return sum(vals.values()) * (n / 1e7), "ncu --set full, profiles/%s (captured at n=1e7, scaled by n)" % \
        os.path.basename(files[-1])


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# CPU arm: the NumPy oracle port of the reference
# ---------------------------------------------------------------------------
def run_oracle(n, ts, outcomes, prior, seed=0):
    """Time the reference's algorithm (NumPy port) on ``len(ts)`` updates of ``n`` particles."""
    import smc_oracle as oracle
    np.random.seed(seed)
    up = oracle.SMCUpdater(oracle.SimplePrecessionModel(), n, FixedPrior(prior))
    step_times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(len(ts)):
            t0 = time.perf_counter()
            up.update(int(outcomes[k]), np.array([ts[k]]))
            step_times.append(time.perf_counter() - t0)
    return step_times, up.resample_count


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args, rank, world):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    # bound the CPU work: ~55 ns per particle-update incl. amortised resamples (BASELINE.md probes)
    budget_s = 150.0
    n = int(min(PARTICLES_PER_GPU, max(10 ** 5, budget_s / ((steps + warm) * 5.5e-8))))
    ts, outcomes = make_data(steps + warm)
    prior = make_prior(n, 99)
    times, n_res = run_oracle(n, ts, outcomes, prior)
    timed = times[warm:]
    total = float(np.sum(timed))
    value = n * steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": "NumPy oracle port of qinfer.SMCUpdater+LiuWestResampler; each step = one update "
                                   "of a %d-particle sample (of the 10^7 workload), %d steps, %d resamples; NumPy "
                                   "ufunc loops are single-threaded, BLAS dot may use %d threads"
                                   % (n, steps, n_res, cpu_threads())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_per_gpu, world):
    return {"workload": "C2 SimplePrecessionModel, %d particles/GPU x %d GPU, t_k=(9/8)^(k mod 100), "
                        "LiuWest a=0.98, resample_thresh 0.5" % (n_per_gpu, world),
            "particles_per_gpu": n_per_gpu, "n_modelparams": 1,
            "l2_policy": "inputs larger than L2: each update streams 240 MB (x, w in, w out) through a 126 MB L2",
            "resampler": "rng=philox (device), scan=fast, draw=auto (guided below 3.2e7 particles, merge above)", "updater": "lazy=True (speculative launch pipelining)"}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def gpu_arm(args, rank, world, local_rank):
    import torch
    import qinfer_b200 as qb
    from qinfer_b200 import _lib

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    n = args.particles
    steps, warm = args.steps, args.warmup
    ts, outcomes = make_data(steps + warm)
    prior = make_prior(n, 99 + rank)

    def new_updater(fuse=None):
        fuse = args.fuse if fuse is None else fuse
        res = qb.LiuWestResampler(a=0.98, rng='philox', seed=1000 + rank, scan='fast')
        if world > 1:
            from qinfer_b200.sharded import ShardedSMCUpdater
            return ShardedSMCUpdater(qb.SimplePrecessionModel(), n * world, FixedPrior(prior), resampler=res, lazy=True,
                                     fuse=fuse)
        return qb.SMCUpdater(qb.SimplePrecessionModel(), n, FixedPrior(prior), resampler=res, lazy=True, fuse=fuse)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # process warm-up (not a step of the workload): run a small cloud through updates AND resamples once so that
        # every kernel of the path is loaded (CUDA loads modules lazily at first launch) before anything is timed
        nsmall = 65536
        if world == 1:
            small = qb.SMCUpdater(qb.SimplePrecessionModel(), nsmall, FixedPrior(prior[:nsmall]), lazy=True,
                                  resampler=qb.LiuWestResampler(a=0.98, rng='philox', seed=5, scan='fast'))
        else:
            from qinfer_b200.sharded import ShardedSMCUpdater
            small = ShardedSMCUpdater(qb.SimplePrecessionModel(), nsmall * world, FixedPrior(prior[:nsmall]), lazy=True,
                                      resampler=qb.LiuWestResampler(a=0.98, rng='philox', seed=5, scan='fast'))
        wts_, wout_ = make_data(30, seed=7)              # its own data: independent of --steps/--warmup
        for k in range(30):
            small.update(int(wout_[k]), wts_[k:k + 1])
        small.est_mean()
        # ... and page-lock the host staging blocks the posterior read-back will recycle (torch's caching host
        # allocator keeps them), as a long-lived process would have done on its first read
        # (the e2e pass reads its prior from page-locked host memory, as the base contract asks; allocated first so
        # that it does not take one of the recycled staging blocks)
        pinned_prior = torch.empty((n, 1), dtype=torch.float64, pin_memory=True)
        pinned_prior.numpy()[:] = prior
        stage = [torch.empty((n, 1), dtype=torch.float64, pin_memory=True),
                 torch.empty((n,), dtype=torch.float64, pin_memory=True)]
        del stage
        if world > 1:
            small.close()
        del small
        # ---------------- value: state resident in HBM ----------------
        up = new_updater()
        for k in range(warm):
            up.update(int(outcomes[k]), ts[k:k + 1])
        cloud = up._cloud
        cloud.preallocate_resample()
        cloud.resample_events = []
        launches0 = cloud.launches
        upd_launches0 = cloud.update_launches
        res0 = up.resample_count
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.35)                             # let nvidia-smi take its first samples
        barrier()
        gc.collect()
        gc.disable()                                     # no collector pauses inside the timed regions (host hygiene)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for k in range(warm, warm + steps):
            up.update(int(outcomes[k]), ts[k:k + 1])
        up._flush()                                      # settle the last (lazily pending) step
        stop.record()
        barrier()
        elapsed_ms = start.elapsed_time(stop)
        clocks = sampler.stop() if rank == 0 else None
        launches = cloud.launches - launches0
        upd_launches = cloud.update_launches - upd_launches0
        n_resamples = up.resample_count - res0
        # average fused-update launch: the timed region minus the resamples (event pairs around each), over K launches.
        # Per-launch event pairs are avoided on purpose: an event between two launches breaks their programmatic
        # dependent-launch overlap.  Gaps between kernels are therefore charged to the kernel (conservative).
        resample_events = list(up._cloud.resample_events)
        resample_ms = float(sum(a.elapsed_time(b) for a, b in resample_events))
        kern_ms = (elapsed_ms - resample_ms) / max(upd_launches, 1)
        posterior_mean = float(up.est_mean()[0])

        # ---------------- e2e: from host arrays, through the public API ----------------
        if world > 1:
            up.close()
        del up, cloud                                    # (a live reference would keep ~1 GB of device buffers
        gc.collect()                                     #  allocated and make the e2e pass cudaMalloc its own)
        torch.cuda.synchronize()
        prior = pinned_prior.numpy()                     # the user's host array, page-locked (allocated above)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        up = new_updater()                               # H2D of the n x 1 prior sample happens here
        up._cloud.preallocate_resample()
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        for k in range(warm, warm + steps):
            up.update(int(outcomes[k]), ts[k:k + 1])
        up._flush()
        torch.cuda.synchronize()
        t_steps = time.perf_counter() - t0 - t_setup
        locs = up.particle_locations                     # D2H posterior
        wts = up.particle_weights
        mean = up.est_mean()
        e1.record()
        barrier()
        t_total = time.perf_counter() - t0
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * t_total)
        e2e_phases = {"setup_and_h2d_ms": 1e3 * t_setup, "updates_ms": 1e3 * t_steps,
                      "readback_ms": 1e3 * (t_total - t_setup - t_steps)}
        assert locs.shape[0] == n and wts.shape[0] == n and np.isfinite(mean).all()
        h2d = (n * 8 + 64 * steps) / steps               # prior upload amortised + per-step experiment record
        d2h = (2 * n * 8 + 8) / steps + 16 * 8           # posterior read-back amortised + per-step stats block

        # ---------------- extra: the same K updates, 8 fused per launch ----------------
        fused = None
        if world == 1 and args.fuse == 1:
            del up
            up = new_updater(fuse=8)
            up._cloud.preallocate_resample()
            for k in range(warm):
                up.update(int(outcomes[k]), ts[k:k + 1])
            up._flush()
            l0 = up._cloud.update_launches
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for k in range(warm, warm + steps):
                up.update(int(outcomes[k]), ts[k:k + 1])
            up._flush()
            f1.record()
            barrier()
            fused_ms = f0.elapsed_time(f1)
            fused = {"updates_per_launch_max": 8, "value": n * steps / (fused_ms * 1e-3), "unit": UNIT,
                     "ms_per_step": fused_ms / steps, "update_launches": up._cloud.update_launches - l0,
                     "resamples": up.resample_count,
                     "note": "same workload and semantics (per-step n_ess check, speculative + roll-back); the fused "
                             "kernel is fp64-pipe bound, not HBM bound"}
        h2d = (n * 8 + 64 * steps) / steps               # prior upload amortised + per-step experiment record
        d2h = (2 * n * 8 + 8) / steps + 16 * 8           # posterior read-back amortised + per-step stats block

    gc.enable()
    if dist is not None:
        t = torch.tensor([elapsed_ms, e2e_ms, kern_ms, resample_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, kern_ms, resample_ms = [float(v) for v in t.cpu()]
        tl = torch.tensor([launches], dtype=torch.int64, device='cuda')
        dist.all_reduce(tl)
        launches = int(tl.item())

    if rank == 0:
        total_particles = n * world
        value = total_particles * steps / (elapsed_ms * 1e-3)
        e2e = total_particles * steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peak_gbs()
        algo_bytes = 8.0 * (1 + 2) * n                   # per launch, per GPU: 8(d+2) B/particle, d = 1
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(n)
        res_each = [a.elapsed_time(b) for a, b in resample_events]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "phases": e2e_phases},
            "gpu_launches": launches, "resamples_in_timed_region": n_resamples,
            "roofline": {"bound": "hbm", "kernel": "fused_update_kernel<PRECESSION>", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_launch": algo_bytes, "avg_launch_ms": kern_ms,
                         "how": "(timed region - sum of event-timed resamples) / fused-update launches",
                         "update_launches": upd_launches, "updates_per_launch": steps / max(upd_launches, 1),
                         "resample_ms_total": resample_ms,
                         "resample_ms_each": [round(v, 3) for v in res_each]},
            "clocks": clocks, "posterior_mean": posterior_mean,
        }
        if res_each:
            # whole Liu-West resample (moments, CDF + guide, fused draw+move, weights, host sqrtm and reads):
            # algorithmic 8(3d+5) B per resampled particle, d = 1
            rb = 8.0 * (3 * 1 + 5) * n
            med = float(np.median(res_each))
            line["resample_roofline"] = {"bound": "hbm", "bytes_per_resample": rb, "median_ms": med,
                                         "achieved": rb / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                         "frac": rb / (med * 1e-3) / 1e9 / peak,
                                         "note": "event-timed around resample(), host work and syncs included"}
        if fused is not None:
            line["fused_f1"] = fused
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
        print(json.dumps(line))
    if dist is not None:
        try:
            up.close()
        except Exception:
            pass
        dist.destroy_process_group()


def cpu_baseline(n):
    """Oracle on the first 14 updates of the same workload at the full particle count (incl. the
    resamples they trigger): ~15-25 s of single-process NumPy."""
    m = 14
    ts, outcomes = make_data(m)
    prior = make_prior(n, 99)
    times, n_res = run_oracle(n, ts, outcomes, prior)
    total = float(np.sum(times))
    return {"value": n * m / total, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": "NumPy oracle port, first %d updates of the schedule at N=%d incl. %d resamples, %.1f s"
                      % (m, n, n_res, total)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=PARTICLES_PER_GPU, help="particles per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fuse", type=int, default=1,
                    help="updates fused per launch in the timed regions (1 = one launch per update, the headline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
