"""bench.py's output contract, checked on the arm that runs without a GPU (``--impl reference``: the NumPy port of the
reference timed on the host cores) and on the helpers the GPU arm uses to fill ``roofline``."""
import json

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_updates_per_sec" and d["unit"] == "particle-updates/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_rank_zero_only_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_helpers():
    sys.path.insert(0, ROOT)
    import bench
    peak, src = bench.measured_peak_gbs()
    assert peak > 1000 and src
    traffic, tsrc = bench.ncu_traffic(10 ** 7)
    assert traffic is None or (1e8 < traffic < 3e8 and "profiles/" in tsrc)    # DRAM bytes of one update launch
    ts, outcomes = bench.make_data(10)
    assert ts.shape == (10,) and set(outcomes.tolist()) <= {0, 1}


# ---------------------------------------------------------------------------
# The GPU arm's data / indexing / record-building logic, driven on CPU: bench.CudaBackend swapped for a double that
# builds the product's SMCUpdater over tests/fake_cloud.py (NumPy stand-in for the device cloud) with the oracle's
# NumPy resampler through the updater's foreign-resampler path.  This is the test that would have caught round 1's
# IndexError (a hard-coded 30-step warm-up indexing a (steps + warmup)-long array).
# ---------------------------------------------------------------------------
class _FakeBackend(object):
    world = 1

    def __init__(self):
        import time
        self.time = time
        self.made = []

    def barrier(self):
        pass

    sync = barrier

    def new_updater(self, n, prior, mode='throughput', fuse=1, seed=1000, sharded=None, lazy=True):
        import smc_oracle as oracle
        from fake_cloud import FakeCloud
        from qinfer_b200.smc import SMCUpdater
        sys.path.insert(0, ROOT)
        import bench

        class HostOnly(SMCUpdater):
            def _rebuild_cloud(self, n_):
                self._cloud = FakeCloud(self._desc, n_)
                self._host_locs = self._host_weights = None

        assert prior.shape == (n, 1)
        up = HostOnly(oracle.SimplePrecessionModel(), n, bench.FixedPrior(prior), lazy=lazy, fuse=fuse,
                      resampler=oracle.LiuWestResampler())
        self.made.append((n, mode, fuse))
        return up

    def pinned(self, array):
        class P(object):
            def __init__(self, a):
                self.a = a.copy()

            def numpy(self):
                return self.a
        return P(array)

    def timer(self):
        time = self.time

        class T(object):
            def start(self):
                self.t0 = time.perf_counter()

            def stop(self):
                self.t1 = time.perf_counter()

            def ms(self):
                return 1e3 * (self.t1 - self.t0)
        return T()

    def collect_resample_events(self, up):
        pass

    def resample_ms(self, up):
        return []

    def launch_counts(self, up):
        return up._cloud.launches, up._cloud.update_launches

    def close(self, up):
        pass

    def reduce(self, values, op="max"):
        return list(values)

    def finish(self):
        pass


@pytest.mark.parametrize("steps,warmup", [(20, 5), (1, 0), (1000, 10)])
def test_gpu_arm_data_path_for_the_drivers_step_counts(steps, warmup, capsys, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    class Args(object):
        pass
    args = Args()
    args.particles, args.steps, args.warmup = 1500, steps, max(warmup, 3)
    args.fuse, args.no_extras, args.no_north_star, args.no_cpu_baseline = 1, False, True, False
    be = _FakeBackend()
    line = bench.gpu_arm(args, 0, 1, 0, backend=be)
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1 and json.loads(out[0])["steps"] == steps
    d = json.loads(out[0])
    assert d["metric"] == "particle_updates_per_sec" and d["n_gpus"] == 1 and d["warmup"] == max(warmup, 3)
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] >= steps
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"])
    assert rf["update_launches"] >= steps and "traffic" in rf and "dram_frac" in rf
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert isinstance(d["cpu_baseline"]["host"], dict) and "numpy" in d["cpu_baseline"]["host"]
    # the double resamples with the oracle's NumPy resampler under the same seed: the parity check is exact
    chk = d["check"]
    assert chk["ok"] and chk["parity_mean_rel_err"] < 1e-9 and chk["parity_resample_count"][0] == \
        chk["parity_resample_count"][1]
    assert d["parity_mode"]["value"] > 0 and d["fused_f1"]["updates_per_launch_max"] == 8
    assert "north_star_1e8" not in d                      # only at the full particle count
    modes = [m for _, m, _ in be.made]
    assert modes.count('parity') >= 2 and 'throughput' in modes
    assert line["posterior_mean"] == d["posterior_mean"]


def test_reference_arm_describes_its_own_algorithm():
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.reference_config(12345, 1)
    assert "MT19937" in cfg["resampler"] and "philox" not in cfg["resampler"].lower()
    assert "lazy" not in cfg["updater"] and cfg["workload"] == bench.workload_config(bench.PARTICLES_PER_GPU, 1)["workload"]


def test_a_failing_extra_does_not_cost_the_headline_line(capsys):
    """An exception inside a sub-record (here: the parity-mode pass) becomes {"ok": false, "error": ...}; the headline
    JSON line with value / e2e / roofline is still printed and the process does not fail."""
    sys.path.insert(0, ROOT)
    import bench

    class Flaky(_FakeBackend):
        def new_updater(self, n, prior, mode='throughput', **kw):
            if mode == 'parity' and kw.get('seed') == 1000:        # the parity-mode PASS (the warm-up uses seed 5)
                raise RuntimeError("simulated failure of an extra")
            return _FakeBackend.new_updater(self, n, prior, mode=mode, **kw)

    class Args(object):
        pass
    args = Args()
    args.particles, args.steps, args.warmup = 1200, 20, 5
    args.fuse, args.no_extras, args.no_north_star, args.no_cpu_baseline = 1, False, True, True
    line = bench.gpu_arm(args, 0, 1, 0, backend=Flaky())
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["roofline"]["frac"] > 0
    assert d["parity_mode"]["ok"] is False and "simulated failure" in d["parity_mode"]["error"]
    assert d["fused_f1"]["value"] > 0                        # the other extras still ran
    assert line["value"] == d["value"]
