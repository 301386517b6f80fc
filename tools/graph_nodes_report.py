#!/usr/bin/env python
"""Summarise an `ncu --graph-profiling node --cache-control none --metrics gpu__time_duration.sum --csv` capture of
tools/profile_clip.py: per-kernel totals plus the node-by-node listing of the second extract unit and the decode unit
that follows it.   usage: tools/graph_nodes_report.py launches.csv > profiles/rNN_graph_nodes_warm.txt"""
import csv
import re
import subprocess
import sys

path = sys.argv[1]
print("# ncu --graph-profiling node --cache-control none --clock-control none --metrics gpu__time_duration.sum")
print("#   python tools/profile_clip.py --frames 16 --global-frames 8 --no-streams --frames-per-stream 8")
print("# Per-node durations of the captured CUDA graphs, caches warm (no flush between kernels), kernels serialised by")
print("# ncu (no PDL overlap).  Units of the profiled clip: extract(16 frames: 8 local + 8 global), decode(8),")
print("# extract(8), decode(8).")
print()
print(subprocess.run([sys.executable, "tools/launch_summary.py", path], capture_output=True, text=True).stdout)
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
rows = [(row[ix["Kernel Name"]], row[ix["Grid Size"]], float(row[ix["Metric Value"]]) / 1000.0) for row in r]
pp = [i for i, x in enumerate(rows) if "preprocess" in x[0]]
s = pp[1]


def short(n):
    n = re.sub(r"dvid::(<unnamed>::)?", "", n)
    return re.sub(r"\(.*", "", n).replace("void ", "")[:60]


print("## second extract unit (8 new frames, 608x1024): node order, grid, us")
tot = 0.0
end = s
for i in range(s, len(rows)):
    n, g, d = rows[i]
    print("%4d  %-36s %-16s %7.1f" % (i - s, short(n), g, d))
    tot += d
    if "topk_mask" in n:
        end = i
        break
print("extract unit total (serialised): %.1f us" % tot)
print()
print("## following decode unit (8 frames, T=4): node order, grid, us")
tot = 0.0
for i in range(end + 1, len(rows)):
    n, g, d = rows[i]
    print("%4d  %-36s %-16s %7.1f" % (i - end - 1, short(n), g, d))
    tot += d
print("decode unit total (serialised): %.1f us" % tot)
