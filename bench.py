"""Headline benchmark: particle-updates/s of the SMC Bayes-update + Liu-West path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], "C2"): SimplePrecessionModel, 10^7 particles per
GPU, prior U[0,1], sequential updates with the exp-sparse schedule t_k = (9/8)^(k mod 100),
true omega = 0.5, default LiuWestResampler(a=0.98), resample_thresh = 0.5.  A "step" is
one ``SMCUpdater.update`` call (fused likelihood/weight/normalisation/n_ess kernel, plus the
whole Liu-West resample whenever n_ess < N/2 — 10 to 20 times per 1000 steps).

Metric: particle-updates/s = particles x steps / time, reference convention of
qinfer/perf_testing.py:250-251 (only ``update`` is timed; prior sampling is not).

  value  inputs resident in HBM, timed with CUDA events around the K update calls
         (throughput mode: device Philox RNG, binned multinomial draw, lazy host settlement).
  e2e    the same K updates through the public API starting from HOST arrays: the
         host->device upload of the prior sample (page-locked), the per-step scalar
         exchanges (experiment record in, normalisation/n_ess out) and the device->host read
         of the posterior mean are inside the timed region; the read-back of the whole
         posterior cloud is timed too and reported as a phase (`with_posterior_readback`).
  roofline  fused-update kernel only: algorithmic bytes 8(d+2) per particle
         over its average device duration; `traffic` = DRAM bytes of one launch from the newest
         committed `ncu --set full` capture (named), `dram_frac` = that traffic over the same time.
  parity_mode  the same K steps with rng='mt19937', scan='exact' (resample indices bit-identical
         to the reference under a legacy seed): what the bit-exact guarantee costs.
  north_star_1e8  N = 10^8 particles on one GPU, 200 steps (BASELINE.json north_star's target).
  c5     (8 GPUs) the same K steps at 1.25e7 particles per GPU = N = 10^8 over the box.
  check  in-run correctness: the parity-mode updater against the CPU baseline's posterior on the
         shared prefix of the schedule (1e-6), the throughput-mode updater within 5 standard errors;
         sharded_check (N > 1): sharded cloud against a single-GPU run of the same cloud.
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
C5_PARTICLES_PER_GPU = 10 ** 8 // 8
TRUE_OMEGA = 0.5
PREFIX_STEPS = 14            # the CPU baseline's sample: the first updates of the schedule at the full particle count
PROCESS_WARMUP_STEPS = 30


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


def drive(up, ts, outcomes, lo, hi):
    """Steps lo .. hi-1 of the workload through the public API (one ``update`` per datum)."""
    if os.environ.get("QB_BENCH_TRACE") and hi - lo <= 64:       # diagnostics: host time after every call
        t0 = time.perf_counter()
        stamps = []
        for k in range(lo, hi):
            up.update(int(outcomes[k]), ts[k:k + 1])
            stamps.append(int((time.perf_counter() - t0) * 1e6))
        sys.stderr.write("[trace rank %s] steps %d..%d host us: %s\n" % (os.environ.get("RANK", "0"), lo, hi, stamps))
        return
    for k in range(lo, hi):
        up.update(int(outcomes[k]), ts[k:k + 1])


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    PERIOD_MS = 200     # the profiling recipe's period: a faster poll holds the driver lock often enough to delay launches

    def __init__(self, gpu_index):
        self.path = "/tmp/qb_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout_s=8.0):
        """Block until nvidia-smi has written its first line: its start-up (NVML initialisation of every GPU of the
        box, hundreds of ms on an 8-GPU node) must not overlap the timed region — it stalls kernel launches."""
        if self.proc is None:
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    time.sleep(0.05)
                    return
            except OSError:
                pass
            time.sleep(0.02)

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
    profiles/ (captured at 10^7 particles; scaled linearly to n), or (None, why).  The capture is named in the
    output so that a reader can tell which build it describes."""
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
    return sum(vals.values()) * (n / 1e7), "ncu --set full capture profiles/%s (n=1e7, flushed L2, scaled by n); not " \
        "re-measured in this run" % os.path.basename(files[-1])


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
    return step_times, up


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_info():
    """CPU model, library versions and BLAS threads of the box the CPU baseline ran on.  Best effort:
    never raises."""
    info = {}
    try:
        info["os_cpu_count"] = os.cpu_count()
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    info["cpu_model"] = ln.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    try:
        import scipy
        info["numpy"], info["scipy"] = np.__version__, scipy.__version__
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_info
        info["blas_threads"] = sorted({int(d.get("num_threads", 0)) for d in threadpool_info()
                                       if d.get("user_api") == "blas"})
    except Exception:
        pass
    return info


def workload_text(n_per_gpu, world):
    return ("C2 SimplePrecessionModel, %d particles/GPU x %d GPU, t_k=(9/8)^(k mod 100), LiuWest a=0.98, "
            "resample_thresh 0.5" % (n_per_gpu, world))


def workload_config(n_per_gpu, world):
    return {"workload": workload_text(n_per_gpu, world),
            "particles_per_gpu": n_per_gpu, "n_modelparams": 1,
            "l2_policy": "inputs larger than L2: each update streams 240 MB (x, w in, w out) through a 126 MB L2",
            "resampler": "rng=philox (device), scan=fast, draw=binned (multinomial counts per bin of 2048 + in-bin "
                         "draws from a shared-memory CDF)",
            "updater": "lazy=True (speculative launch pipelining)",
            "warm_up": "process warm-up at full size, then 25 ms of SM spin-up (the sampler start-up / gc / barrier "
                       "before a timed region leave the GPU idle and its clocks low), then the W warm-up steps"}


def reference_config(n, world):
    return {"workload": workload_text(PARTICLES_PER_GPU, world),
            "particles_per_gpu": PARTICLES_PER_GPU, "particles_timed": n, "n_modelparams": 1,
            "resampler": "the reference's: np.random (legacy MT19937) uniforms + np.cumsum + searchsorted + randn, host",
            "updater": "the reference's call-by-call SMCUpdater.update (NumPy port, single process)"}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    # bound the CPU work: ~55 ns per particle-update incl. amortised resamples (BASELINE.md probes)
    budget_s = 150.0
    n = int(min(PARTICLES_PER_GPU, max(10 ** 5, budget_s / (max(steps + warm, 1) * 5.5e-8))))
    ts, outcomes = make_data(steps + warm)
    prior = make_prior(n, 99)
    times, up = run_oracle(n, ts, outcomes, prior)
    timed = times[warm:]
    total = float(np.sum(timed))
    value = n * steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": reference_config(n, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": "NumPy oracle port of qinfer.SMCUpdater+LiuWestResampler; each step = one update "
                                   "of a %d-particle sample (of the 10^7 workload), %d steps, %d resamples; NumPy "
                                   "ufunc loops are single-threaded, BLAS dot may use %d threads"
                                   % (n, steps, up.resample_count, cpu_threads()),
                         "host": host_info()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(n):
    """Oracle on the first PREFIX_STEPS updates of the same workload at the full particle count (incl. the
    resamples they trigger): ~15-25 s of single-process NumPy.  Also returns the posterior the GPU arm checks
    itself against."""
    m = PREFIX_STEPS
    ts, outcomes = make_data(m)
    prior = make_prior(n, 99)
    times, up = run_oracle(n, ts, outcomes, prior)
    total = float(np.sum(times))
    rec = {"value": n * m / total, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
           "sample": "NumPy oracle port, first %d updates of the schedule at N=%d incl. %d resamples, %.1f s"
                     % (m, n, up.resample_count, total),
           "host": host_info()}
    post = {"mean": float(up.est_mean()[0]), "cov": float(up.est_covariance_mtx()[0, 0]), "n_ess": float(up.n_ess),
            "resample_count": int(up.resample_count),
            "normalization_record": [float(np.ravel(v)[0]) for v in up.normalization_record]}
    return rec, post


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class CudaBackend(object):
    """Everything of the GPU arm that touches torch / the engine, behind a small surface so that the arm's data and
    indexing logic can be driven on CPU by a test double (tests/test_bench_contract.py)."""

    def __init__(self, local_rank, world):
        import torch
        import qinfer_b200 as qb
        self.torch, self.qb, self.world = torch, qb, world
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl")
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def sync(self):
        self.torch.cuda.synchronize()

    def spin_up(self, ms=25.0):
        """Keep the SMs busy for ``ms`` on a scratch buffer.  The host-side preparations of a timed region (clock
        sampler start-up, garbage collection, barrier) leave the GPU idle for hundreds of milliseconds, its clocks
        fall back to idle, and the first ~millisecond afterwards runs below the clocks of a busy device — comparable
        to a whole 20-step region.  This is process warm-up; the W warm-up steps follow it."""
        torch = self.torch
        buf = getattr(self, "_spin_buf", None)
        if buf is None:
            buf = self._spin_buf = torch.ones((1 << 24,), dtype=torch.float64, device="cuda")
        t_end = time.perf_counter() + ms * 1e-3
        while time.perf_counter() < t_end:
            for _ in range(8):
                buf.mul_(1.0)
            torch.cuda.synchronize()

    def resampler(self, mode, seed):
        if mode == 'parity':
            return self.qb.LiuWestResampler(a=0.98, rng='mt19937', scan='exact')
        return self.qb.LiuWestResampler(a=0.98, rng='philox', seed=seed, scan='fast')

    def new_updater(self, n, prior, mode='throughput', fuse=1, seed=1000, sharded=None, lazy=True):
        qb = self.qb
        sharded = (self.world > 1) if sharded is None else sharded
        res = self.resampler(mode, seed)
        if sharded:
            from qinfer_b200.sharded import ShardedSMCUpdater
            up = ShardedSMCUpdater(qb.SimplePrecessionModel(), n * self.world, FixedPrior(prior), resampler=res,
                                   lazy=lazy, fuse=fuse)
        else:
            up = qb.SMCUpdater(qb.SimplePrecessionModel(), n, FixedPrior(prior), resampler=res, lazy=lazy, fuse=fuse)
        cloud = up._cloud
        if mode == 'parity':
            cloud.preallocate_resample()
        else:
            cloud.preallocate_binned()
        if os.environ.get("QB_BENCH_TRACE"):                 # diagnostics: host time of every stage of a resample
            stages = []

            def wrap(obj, name):
                fn = getattr(obj, name, None)
                if fn is None:
                    return

                def inner(*a, **k):
                    t_a = time.perf_counter()
                    r = fn(*a, **k)
                    stages.append("%s %d" % (name, (time.perf_counter() - t_a) * 1e6))
                    if name in ("adopt_binned", "_finish_resample"):
                        sys.stderr.write("[trace rank %s] resample stages (us): %s\n"
                                         % (os.environ.get("RANK", "0"), ", ".join(stages)))
                        del stages[:]
                    return r
                setattr(obj, name, inner)
            for nm in ("binned_sums", "binned_count", "binned_move", "binned_resample", "binned_moments_wait",
                       "binned_retry_wait", "binned_counters_wait", "_alt_slab", "adopt_binned"):
                wrap(cloud, nm)
            if sharded:
                wrap(up._comm, "all_gather_rows")
                wrap(up, "_liu_west_consts")
        return up

    def parity_js(self, up):
        """Global parent indices of the last parity-mode resample (this rank's slab of them when sharded)."""
        js = getattr(up, "last_parity_js", None)
        if js is None:
            js = up._cloud._js
        return None if js is None else js.cpu().numpy()

    def pinned(self, array):
        t = self.torch.empty(array.shape, dtype=self.torch.float64, pin_memory=True)
        t.numpy()[:] = array
        return t

    def timer(self):
        torch = self.torch

        class T(object):
            def start(self):
                self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.a.record()

            def stop(self):
                self.b.record()

            def ms(self):
                torch.cuda.synchronize()
                return self.a.elapsed_time(self.b)
        return T()

    def collect_resample_events(self, up):
        up._cloud.resample_events = []

    def resample_ms(self, up):
        return [a.elapsed_time(b) for a, b in up._cloud.resample_events]

    def launch_counts(self, up):
        return up._cloud.launches, up._cloud.update_launches

    def close(self, up):
        if hasattr(up, 'close'):
            try:
                up.close()
            except Exception:
                pass

    def reduce(self, values, op="max"):
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def process_warmup(be, prior):
    """Not a step of the workload: run a small cloud through updates AND resamples once so that every kernel of
    the path is loaded (CUDA loads modules lazily at first launch) before anything is timed.  Uses its own data,
    independent of --steps / --warmup."""
    nsmall = min(65536, prior.shape[0])
    wts, wout = make_data(PROCESS_WARMUP_STEPS, seed=7)
    # a small cloud in both modes (module loading), then a throw-away cloud of the workload's size in the timed mode:
    # the allocator, the pinned staging blocks and (sharded) NCCL's channels for these message sizes exist before
    # the timed updater is built
    plan = [(nsmall, 'throughput')] + ([(nsmall, 'parity')] if be.world == 1 else []) + [(prior.shape[0], 'throughput')] \
        + ([(prior.shape[0], 'parity')] if be.world == 1 else [])    # (the device MT19937 path's first call at a new size
                                                                     #  costs ~0.1 s once per process)
    for nsz, mode in plan:
        small = be.new_updater(nsz, prior[:nsz], mode=mode, seed=5)
        if mode == 'parity':
            np.random.seed(123)
        drive(small, wts, wout, 0, PROCESS_WARMUP_STEPS)
        small.est_mean()
        be.close(small)
        del small


def timed_run(be, n, prior, ts, outcomes, warm, steps, mode='throughput', fuse=1, seed=1000, clocks=None):
    """W untimed warm-up steps, then exactly K timed steps between barriers; device time by CUDA events."""
    up = be.new_updater(n, prior, mode=mode, fuse=fuse, seed=seed)
    if mode == 'parity':
        np.random.seed(0)
    if clocks is not None:
        clocks.start()
        clocks.wait_first_sample()                   # nvidia-smi's start-up stays outside the timed region
    gc.collect()                                     # (before the barrier: a collection takes milliseconds and would
    gc.disable()                                     #  skew the ranks' entry into the timed region)
    if hasattr(be, "spin_up"):
        be.spin_up()                                 # the preparations above left the GPU idle: clocks back up first
    drive(up, ts, outcomes, 0, warm)                 # the W untimed warm-up steps, right before the timed ones
    up._flush()
    be.collect_resample_events(up)
    l0, u0 = be.launch_counts(up)
    r0 = up.resample_count
    be.barrier()
    t = be.timer()
    t.start()
    drive(up, ts, outcomes, warm, warm + steps)
    up._flush()                                      # settle the last (lazily pending) step
    t.stop()
    be.barrier()
    gc.enable()
    elapsed_ms = t.ms()
    clk = clocks.stop() if clocks is not None else None
    l1, u1 = be.launch_counts(up)
    res_each = be.resample_ms(up)
    out = {"elapsed_ms": elapsed_ms, "launches": l1 - l0, "update_launches": u1 - u0,
           "resamples": up.resample_count - r0, "resample_ms_each": res_each, "clocks": clk,
           "posterior_mean": float(up.est_mean()[0]), "n_ess": float(up.n_ess)}
    be.close(up)
    del up
    gc.collect()
    return out


def e2e_run(be, n, pinned_prior, ts, outcomes, warm, steps, fuse=1, seed=1000):
    """From HOST arrays through the public API: prior upload (page-locked source), K updates, posterior mean read;
    then the read-back of the whole posterior cloud, timed separately."""
    prior = pinned_prior.numpy()
    gc.collect()
    if hasattr(be, "spin_up"):
        be.spin_up()
    be.barrier()
    t0 = time.perf_counter()
    t = be.timer()
    t.start()
    up = be.new_updater(n, prior, fuse=fuse, seed=seed)    # H2D of the n x 1 prior sample happens here
    be.sync()
    t_setup = time.perf_counter() - t0
    drive(up, ts, outcomes, warm, warm + steps)
    up._flush()
    mean = up.est_mean()                                    # D2H of the result the caller asks for
    t.stop()
    t_core = time.perf_counter() - t0
    locs = up.particle_locations                            # D2H of the whole posterior (a user who plots it)
    wts = up.particle_weights
    t_total = time.perf_counter() - t0
    be.barrier()
    core_ms = max(t.ms(), 1e3 * t_core)
    # (a sharded cloud's slab floats around n after a resample)
    assert locs.shape[0] == wts.shape[0] and abs(locs.shape[0] - n) <= 0.05 * n + 4096 and np.isfinite(mean).all()
    be.close(up)
    del up, locs, wts
    gc.collect()
    return {"core_ms": core_ms, "setup_and_h2d_ms": 1e3 * t_setup, "updates_ms": 1e3 * (t_core - t_setup),
            "posterior_readback_ms": 1e3 * (t_total - t_core), "with_readback_ms": 1e3 * t_total}


def prefix_check(be, n, post):
    """The GPU path against the CPU baseline's posterior after the shared prefix of the schedule, from the same prior:
    parity mode (legacy MT19937 stream on the device + exact scan) must agree to 1e-6 (north_star) and take the same
    number of resamples; throughput mode (Philox, binned draw) must be statistically consistent."""
    ts, outcomes = make_data(PREFIX_STEPS)
    prior = make_prior(n, 99)
    out = {}
    up = be.new_updater(n, prior, mode='parity', fuse=1, sharded=False, lazy=False)
    np.random.seed(0)
    drive(up, ts, outcomes, 0, PREFIX_STEPS)
    mean, cov = float(up.est_mean()[0]), float(up.est_covariance_mtx()[0, 0])
    norm = np.array([float(np.ravel(v)[0]) for v in up.normalization_record])
    out["parity_mean_rel_err"] = abs(mean - post["mean"]) / abs(post["mean"])
    out["parity_cov_rel_err"] = abs(cov - post["cov"]) / abs(post["cov"])
    out["parity_norm_record_max_rel_err"] = float(np.max(np.abs(norm / np.array(post["normalization_record"]) - 1)))
    out["parity_resample_count"] = [int(up.resample_count), post["resample_count"]]
    be.close(up)
    del up
    up = be.new_updater(n, prior, mode='throughput', fuse=1, sharded=False)
    drive(up, ts, outcomes, 0, PREFIX_STEPS)
    mean_t, ess_t = float(up.est_mean()[0]), float(up.n_ess)
    se = np.sqrt(post["cov"] / max(post["n_ess"], 1.0) + post["cov"] / max(ess_t, 1.0))
    out["throughput_mean_sigmas"] = abs(mean_t - post["mean"]) / se
    out["throughput_resample_count"] = [int(up.resample_count), post["resample_count"]]
    be.close(up)
    del up
    out["tolerance"] = "parity: 1e-6 relative on mean and covariance, equal resample count; throughput: 5 standard errors"
    out["ok"] = bool(out["parity_mean_rel_err"] < 1e-6 and out["parity_cov_rel_err"] < 1e-6 and
                     out["parity_norm_record_max_rel_err"] < 1e-9 and
                     out["parity_resample_count"][0] == out["parity_resample_count"][1] and
                     out["throughput_mean_sigmas"] < 5.0)
    return out


def sharded_check(be, rank, world):
    """N > 1: a small sharded cloud against a single-GPU run of the same cloud on rank 0 — records and n_ess up to the
    first resample to 1e-10 (different summation order), posterior mean within 5 standard errors at the end."""
    n_local, m = 131072, 40
    ts, outcomes = make_data(m, seed=77)
    full = make_prior(n_local * world, 4242)
    up = be.new_updater(n_local, full[rank * n_local:(rank + 1) * n_local], seed=31 + rank)
    drive(up, ts, outcomes, 0, m)
    rec = np.array([float(np.ravel(v)[0]) for v in up.normalization_record])
    mean, cov, ess, count = float(up.est_mean()[0]), float(up.est_covariance_mtx()[0, 0]), float(up.n_ess), \
        int(up.resample_count)
    be.close(up)
    del up
    out = None
    if rank == 0:
        one = be.new_updater(n_local * world, full, seed=31, sharded=False)
        first = None
        for k in range(m):
            one.update(int(outcomes[k]), ts[k:k + 1])
            if first is None and one.resample_count > 0:
                first = k
        rec1 = np.array([float(np.ravel(v)[0]) for v in one.normalization_record])
        upto = m if first is None else first + 1          # records up to and including the step that triggered it
        mean1, cov1, ess1 = float(one.est_mean()[0]), float(one.est_covariance_mtx()[0, 0]), float(one.n_ess)
        se = np.sqrt(cov / max(ess, 1.0) + cov1 / max(ess1, 1.0))
        out = {"world": world, "particles": n_local * world, "steps": m,
               "records_before_first_resample": int(upto),
               "norm_record_max_rel_err": float(np.max(np.abs(rec[:upto] / rec1[:upto] - 1))),
               "resample_count": [count, int(one.resample_count)],
               "mean_sigmas": abs(mean - mean1) / se, "mean": [mean, mean1]}
        out["ok"] = bool(out["norm_record_max_rel_err"] < 1e-10 and out["mean_sigmas"] < 5.0 and
                         abs(count - int(one.resample_count)) <= 1)
        del one
    be.barrier()
    # parity mode on the sharded cloud: legacy MT19937 stream + the exact scan chained across the slabs;
    # the global resample indices must equal the single-GPU parity run's
    if hasattr(be, "parity_js"):
        np.random.seed(99)                                 # the same legacy state on every rank
        up = be.new_updater(n_local, full[rank * n_local:(rank + 1) * n_local], mode='parity', lazy=False)
        drive(up, ts, outcomes, 0, m)
        js, count_p, mean_p = be.parity_js(up), int(up.resample_count), float(up.est_mean()[0])
        be.close(up)
        del up
        if rank == 0:
            np.random.seed(99)
            one = be.new_updater(n_local * world, full, mode='parity', sharded=False, lazy=False)
            drive(one, ts, outcomes, 0, m)
            js1 = be.parity_js(one)[:n_local]
            mean1 = float(one.est_mean()[0])
            par = {"resample_count": [count_p, int(one.resample_count)],
                   "last_resample_js_equal_frac": float(np.mean(js == js1)) if js is not None else None,
                   "mean_rel_err": abs(mean_p - mean1) / abs(mean1)}
            par["ok"] = bool(count_p == int(one.resample_count) and count_p > 0 and par["mean_rel_err"] < 1e-6 and
                             par["last_resample_js_equal_frac"] is not None and
                             par["last_resample_js_equal_frac"] >= 0.999)
            out["parity_mode"] = par
            out["ok"] = bool(out["ok"] and par["ok"])
            del one
        be.barrier()
    return out


def roofline_record(run, n, steps):
    peak, peak_src = measured_peak_gbs()
    resample_ms = float(sum(run["resample_ms_each"]))
    # average fused-update launch: the timed region minus the resamples (event pairs around each), over the launches.
    # Per-launch event pairs are avoided on purpose: an event between two launches breaks their programmatic
    # dependent-launch overlap.  Gaps between kernels are therefore charged to the kernel (conservative).
    kern_ms = (run["elapsed_ms"] - resample_ms) / max(run["update_launches"], 1)
    algo_bytes = 8.0 * (1 + 2) * n                   # per launch, per GPU: 8(d+2) B/particle, d = 1
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(n)
    return {"bound": "hbm", "kernel": "fused_update_kernel<PRECESSION>", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "dram_frac": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
            "note": "achieved = ALGORITHMIC bytes over time; consecutive launches walk the slab in opposite "
                    "directions, so part of it is served by the 126 MB L2 and DRAM moves less (traffic); dram_frac "
                    "is the DRAM share of the measured copy peak",
            "peak_source": peak_src, "bytes_per_launch": algo_bytes, "avg_launch_ms": kern_ms,
            "how": "(timed region - sum of event-timed resamples) / fused-update launches",
            "update_launches": run["update_launches"], "updates_per_launch": steps / max(run["update_launches"], 1),
            "resample_ms_total": resample_ms, "resample_ms_each": [round(v, 3) for v in run["resample_ms_each"]]}


def resample_roofline(run, n):
    if not run["resample_ms_each"]:
        return None
    peak, _ = measured_peak_gbs()
    # whole Liu-West resample (moments + constants, counts, draw + move + weights, retry launch, host wait):
    # algorithmic 8(3d+5) B per resampled particle, d = 1; the binned path moves 8(3d+3)
    rb = 8.0 * (3 * 1 + 5) * n
    med = float(np.median(run["resample_ms_each"]))
    return {"bound": "hbm", "bytes_per_resample": rb, "median_ms": med, "achieved": rb / (med * 1e-3) / 1e9,
            "peak": peak, "unit": "GB/s", "frac": rb / (med * 1e-3) / 1e9 / peak,
            "note": "event-timed around resample(), host work and syncs included"}


def guarded(name, fn):
    """Run one EXTRA sub-record; an exception becomes {"ok": false, "error": ...} instead of costing the whole line."""
    try:
        return fn()
    except Exception as e:        # noqa: BLE001 - anything: the headline measurement is already in hand
        import traceback
        sys.stderr.write("bench.py: sub-record %s failed:\n%s\n" % (name, traceback.format_exc()))
        return {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}


def gpu_arm(args, rank, world, local_rank, backend=None):
    be = backend if backend is not None else CudaBackend(local_rank, world)
    n = args.particles
    steps, warm = args.steps, args.warmup
    ts, outcomes = make_data(steps + warm)
    prior = make_prior(n, 99 + rank)
    extras = not args.no_extras
    line = None
    fused = parity = star = c5 = shard = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        process_warmup(be, prior)
        # page-lock the host block of the e2e pass (as the base contract asks) and the staging blocks its posterior
        # read-back will recycle (torch's caching host allocator keeps them), as a long-lived process would have
        pinned_prior = be.pinned(prior)
        stage = [be.pinned(prior), be.pinned(prior[:, 0])]
        del stage
        # ---------------- value: state resident in HBM ----------------
        sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("QB_BENCH_NO_CLOCKS")) else None
        run = timed_run(be, n, prior, ts, outcomes, warm, steps, fuse=args.fuse, seed=1000 + rank, clocks=sampler)
        # ---------------- e2e: from host arrays, through the public API ----------------
        e2e = e2e_run(be, n, pinned_prior, ts, outcomes, warm, steps, fuse=args.fuse, seed=1000 + rank)
        rec_bytes = 64                                   # experiment record in (t, outcome, flags); stats block out
        h2d = n * 8.0 / steps + rec_bytes                # prior upload amortised + per-step experiment record
        d2h = 64 + 8.0 / steps                           # per-step stats block + the posterior mean
        # The sub-records below are EXTRAS: whatever happens in one of them, the headline line is still printed
        # (`guarded` turns an exception into an {"ok": false, "error": ...} sub-record).
        def fused_record():
            # the same K updates, 8 fused per launch
            f = timed_run(be, n, prior, ts, outcomes, warm, steps, fuse=8, seed=1000)
            return {"updates_per_launch_max": 8, "value": n * steps / (f["elapsed_ms"] * 1e-3), "unit": UNIT,
                    "ms_per_step": f["elapsed_ms"] / steps, "update_launches": f["update_launches"],
                    "resamples": f["resamples"],
                    "note": "same workload and semantics (per-step n_ess check, speculative + roll-back); the fused "
                            "kernel is fp64-pipe bound, not HBM bound"}

        def parity_record():
            # the bit-exact mode: legacy MT19937 stream continued on the device, exact (np.cumsum-rounding) scan
            p = timed_run(be, n, prior, ts, outcomes, warm, steps, mode='parity', fuse=1)
            return {"resampler": "rng=mt19937 (NumPy's legacy stream, generated on the device), scan=exact, staged "
                                 "draw: resample indices bit-identical to the reference under np.random.seed",
                    "value": n * steps / (p["elapsed_ms"] * 1e-3), "unit": UNIT,
                    "ms_per_step": p["elapsed_ms"] / steps, "resamples": p["resamples"],
                    "resample_ms_each": [round(v, 3) for v in p["resample_ms_each"]],
                    "posterior_mean": p["posterior_mean"]}

        def north_star_record():
            big = 10 ** 8
            bts, bout = make_data(205)
            s = timed_run(be, big, make_prior(big, 7), bts, bout, 5, 200, fuse=1, seed=2000)
            sr = roofline_record(s, big, 200)
            return {"workload": "SimplePrecessionModel, N = 1e8 particles on ONE GPU, 200 updates (north_star)",
                    "value": big * 200 / (s["elapsed_ms"] * 1e-3), "unit": UNIT,
                    "ms_per_step": s["elapsed_ms"] / 200, "resamples": s["resamples"],
                    "resample_ms_each": [round(v, 3) for v in s["resample_ms_each"]],
                    "roofline_frac": sr["frac"], "avg_update_launch_ms": sr["avg_launch_ms"],
                    "target": ">= 1e9 particle-updates/s at >= 60 % of the HBM roofline"}

        def c5_record():
            nc = C5_PARTICLES_PER_GPU
            c = timed_run(be, nc, make_prior(nc, 199 + rank), ts, outcomes, warm, steps, fuse=args.fuse,
                          seed=3000 + rank)
            c_ms = be.reduce([c["elapsed_ms"]])[0]
            return {"workload": "C5: N = 1e8 over 8 GPUs, 1.25e7 particles per GPU, same %d steps" % steps,
                    "value": nc * world * steps / (c_ms * 1e-3), "unit": UNIT, "ms_per_step": c_ms / steps,
                    "resamples": c["resamples"]}

        if extras and world == 1 and args.fuse == 1:
            fused = guarded("fused_f1", fused_record)
            parity = guarded("parity_mode", parity_record)
            if n == PARTICLES_PER_GPU and not args.no_north_star:
                star = guarded("north_star_1e8", north_star_record)
        if extras and world == 8 and n == PARTICLES_PER_GPU:
            c5 = guarded("c5", c5_record)
        if extras and world > 1:
            shard = guarded("sharded_check", lambda: sharded_check(be, rank, world))

    elapsed_ms, core_ms = be.reduce([run["elapsed_ms"], e2e["core_ms"]])
    launches = int(be.reduce([run["launches"]], op="sum")[0])
    if rank == 0:
        total_particles = n * world
        value = total_particles * steps / (elapsed_ms * 1e-3)
        run_max = dict(run, elapsed_ms=elapsed_ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
            "e2e": {"value": total_particles * steps / (core_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "timed": "prior upload from page-locked host memory + K updates (host scalars in, stats out) + "
                             "posterior mean read",
                    "phases": {k: e2e[k] for k in ("setup_and_h2d_ms", "updates_ms", "posterior_readback_ms")},
                    "with_posterior_readback": {"value": total_particles * steps / (e2e["with_readback_ms"] * 1e-3),
                                                "d2h_bytes_per_step": d2h + 16.0 * n / steps}},
            "gpu_launches": launches, "resamples_in_timed_region": run["resamples"],
            "roofline": roofline_record(run_max, n, steps),
            "clocks": run["clocks"], "posterior_mean": run["posterior_mean"],
        }
        rr = resample_roofline(run, n)
        if rr is not None:
            line["resample_roofline"] = rr
        for key, val in (("fused_f1", fused), ("parity_mode", parity), ("north_star_1e8", star), ("c5", c5),
                         ("sharded_check", shard)):
            if val is not None:
                line[key] = val
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], post = cpu_baseline(n)
            if extras:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    line["check"] = guarded("check", lambda: prefix_check(be, n, post))
        print(json.dumps(line))
        sys.stdout.flush()
        bad = [k for k in ("check", "sharded_check") if k in line and not line[k].get("ok", False)]
        if bad:
            sys.stderr.write("bench.py: in-run correctness check failed: %s\n" % ", ".join(bad))
    be.finish()
    return line


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=PARTICLES_PER_GPU, help="particles per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip fused_f1 / parity_mode / north_star_1e8 / c5 / the correctness checks")
    ap.add_argument("--no-north-star", action="store_true")
    ap.add_argument("--fuse", type=int, default=1,
                    help="updates fused per launch in the timed regions (1 = one launch per update, the headline)")
    args = ap.parse_args(argv)
    if args.steps < 1:
        ap.error("--steps must be at least 1")
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
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
