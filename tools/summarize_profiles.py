"""Turn raw ncu output in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> launches <launches.csv> "<command>"
    python tools/summarize_profiles.py <tag> full <report.ncu-rep> [name]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(tag, path, command):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        a = agg.setdefault(row['Kernel Name'], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["# %s — ncu launch list of `%s`" % (tag, command),
           "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
           "", "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f%% | %.1f |" % (k.split('(')[0].replace('void ', ''), a[0], a[1],
                                                          100 * a[1] / tot, a[1] / a[0]))
    open(os.path.join(OUT, "%s_launches_summary.md" % tag), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(tag, rep, name):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    path = os.path.join(OUT, "%s_%s_ncu_full.csv" % (tag, name))
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(data))])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                w.writerow([h, units[i]] + [r[i] for r in data])
    print(open(path).read())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[2] == "launches":
        launches(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "?")
    else:
        full(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else os.path.basename(sys.argv[3]).split('.')[0])
