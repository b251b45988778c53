#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, mean, share.
usage: summarize_launches.py launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if v != v:
        continue
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[d["Metric Unit"]]
    if "pc::" not in d["Kernel Name"]:  # torch fill / randn launches of the bench set-up are not part of a step
        continue
    key = (d["Kernel Name"][:90], d["Block Size"], d["Grid Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# per-launch device time from ncu (cold-cache, serialised): compare SHARES, not absolutes")
print("%-92s %-14s %-14s %5s %11s %10s %6s" % ("kernel", "block", "grid", "n", "total_us", "mean_us", "share"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-92s %-14s %-14s %5d %11.1f %10.1f %6.3f" % (k[0], k[1], k[2], a[0], a[1], a[1] / a[0], a[1] / tot))
