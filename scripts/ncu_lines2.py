#!/usr/bin/env python
"""Per-CUDA-source-line summary of an .ncu-rep: stall samples and executed warp instructions.
Usage: scripts/ncu_lines2.py <report.ncu-rep> [min_samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
tot_s = tot_i = 0
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or not r[0]:
        continue
    try:
        ns = int(r[hdr.index("# Samples")]); ni = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    tot_s += ns; tot_i += ni
    lines.append((int(r[0]), ns, ni, r[1].strip()))
print(f"total samples {tot_s}, warp instructions {tot_i}")
for ln, ns, ni, src in lines:
    if ns >= mins or ni >= 0.02 * tot_i:
        print(f"{ln:5d} samp {ns:6d} ({100*ns/max(tot_s,1):5.1f}%) inst {ni:10d} ({100*ni/max(tot_i,1):5.1f}%)  {src[:110]}")
