#!/usr/bin/env python
"""Per-instruction memory-pipe cost from an .ncu-rep source page: shared wavefronts, global L1 tag
requests, executed count and stall samples, for every SASS instruction that touches memory or
collects >= 1% of the stall samples.  Usage: scripts/ncu_lines.py <report.ncu-rep>"""
import csv
import io
import subprocess
import sys

src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    data.append(r)
ix = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "L1 Tag Requests Global", "L1 Wavefronts Shared",
                                "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global")}
def num(r, k):
    try:
        return int(r[ix[k]])
    except ValueError:
        return 0
tot_s = sum(num(r, "# Samples") for r in data) or 1
tot_sh = sum(num(r, "L1 Wavefronts Shared") for r in data)
tot_g = sum(num(r, "L1 Tag Requests Global") for r in data)
print(f"total: samples {tot_s}, shared wavefronts {tot_sh}, global L1 tag requests {tot_g}, instructions {sum(num(r, 'Instructions Executed') for r in data)}")
print(f"{'idx':>5s} {'samples%':>8s} {'executed':>9s} {'sh_wf':>9s} {'sh_ideal':>9s} {'g_tag':>9s} {'l2_sect':>9s}  sass")
for i, r in enumerate(data):
    s, sh, g = num(r, "# Samples"), num(r, "L1 Wavefronts Shared"), num(r, "L1 Tag Requests Global")
    if sh or g or s >= 0.01 * tot_s:
        print(f"{i:5d} {100 * s / tot_s:8.1f} {num(r, 'Instructions Executed'):9d} {sh:9d} {num(r, 'L1 Wavefronts Shared Ideal'):9d} {g:9d} "
              f"{num(r, 'L2 Theoretical Sectors Global'):9d}  {r[ix['Source']].strip()[:70]}")
