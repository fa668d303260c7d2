#!/usr/bin/env python
"""Per-kernel shares of ONE refinement step from an ncu launch list (--metrics gpu__time_duration.sum --csv): the window that
starts at the 30th-from-last linearisation = the last dsc_optimize call of the command (bench.py's last timed step).
usage: python profiles/launch_shares.py launches.csv [linearisations per step]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
k, v = h.index("Kernel Name"), h.index("Metric Value")
names = [r[k].split("(")[0].replace("dsc::", "").split("<")[0] for r in rows[1:]]
times = [float(r[v].replace(",", "")) for r in rows[1:]]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lin = [i for i, n in enumerate(names) if "linearize_ell_kernel" in n]
start = lin[-per_step] if len(lin) >= per_step else 0
# the step ends with the write-back (export_kernel) that follows it, if that was captured too
end = next((i for i in range(lin[-1] if lin else 0, len(names)) if "export_kernel" in names[i]), len(names) - 1) + 1
agg = collections.OrderedDict()
for n, t in zip(names[start:end], times[start:end]):
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# launches {start}..{end} of {len(names)}: one step = {per_step} LM iterations; total {tot / 1e3:.1f} us of kernel time")
print(f"{'kernel':32s} {'launches':>8s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:32s} {c:8d} {t / 1e3:12.1f} {t / 1e3 / c:9.1f} {t / tot:7.3f}")
