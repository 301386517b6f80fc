#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count / total / share."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for row in r:
    name = row["Kernel Name"]
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = short[:70]
    key = (short, row["Block Size"])
    ns = float(row["Metric Value"].replace(",", ""))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ns; a[2] = max(a[2], ns)
    tot += ns
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
print("%-72s %-14s %6s %10s %8s %9s %6s" % ("kernel", "block", "n", "total_us", "avg_us", "max_us", "share"))
for (k, b), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %-14s %6d %10.1f %8.2f %9.1f %5.1f%%" % (k, b, a[0], a[1] / 1e3, a[1] / 1e3 / a[0], a[2] / 1e3, 100 * a[1] / tot))
