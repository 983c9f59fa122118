#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:
launch count, total / mean device time and share of the captured time.
Usage: scripts/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_<tag>.txt"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = OrderedDict()
total = 0.0
n = 0
for r in rd:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"^.*?::", "", r[ki])
    name = re.sub(r"\(.*$", "", name)
    ns = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += ns
    total += ns
    n += 1
print(f"# {path}: {n} launches, {total / 1e6:.3f} ms of kernel time (ncu per-launch times: cold cache, serialised)")
print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>10s} {'mean_us':>9s} {'share':>7s}  grid x block (last)")
for name, (cnt, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:48]:48s} {cnt:8d} {ns / 1e6:10.3f} {ns / cnt / 1e3:9.2f} {100 * ns / total:6.1f}%  {g} x {b}")
