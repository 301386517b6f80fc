#!/usr/bin/env python
"""Top stall-sample SASS lines of an ncu source page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (i, 100.0 * s / max(tot, 1), r[ix["Source"]].strip()[:70], ", ".join("%s=%d" % (h[6:], c) for c, h in why if c)))
