"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_shares.py profiles/r1_bench_launches.csv
"""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
by = collections.defaultdict(list)
for r in rows:
    by[r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "")].append(float(r["Metric Value"]))
tot = sum(sum(v) for v in by.values())
print("%d launches, %.1f us of device time (ncu: cold caches, serialised)" % (len(rows), tot / 1e3))
for k, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
    print("%6d launches  %10.1f us  %5.1f%%  avg %8.2f us  %s" % (len(v), sum(v) / 1e3, 100 * sum(v) / tot,
                                                                 sum(v) / len(v) / 1e3, k[:110]))
