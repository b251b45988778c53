#!/usr/bin/env python
"""Launch-by-launch listing of an ncu --csv log (gpu__time_duration.sum): python scripts/print_launches.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, tot = None, 0.0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[d["Metric Unit"]]
    tot += v
    name = d["Kernel Name"].replace("void ", "").replace("pc::<unnamed>::", "").replace("unnamed>::", "")
    print("%-64s %-14s %9.1f us" % (name[:64], d["Grid Size"], v))
print("total %.1f us" % tot)
