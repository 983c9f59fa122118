#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics and the hottest SASS
instructions by warp-stall samples.  Usage: scripts/ncu_summary.py <report.ncu-rep> [top_frac]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("=== ", r[hdr.index("Kernel Name")][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:80s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
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
    try:
        data.append((int(r[hdr.index("# Samples")]), r[hdr.index("Source")].strip(), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print(f"--- first kernel: {tot} stall samples over {len(data)} SASS instructions; hottest (>= {100*thresh:.1f}%):")
names = ["stall_long_sb", "stall_barrier", "stall_short_sb", "stall_mio", "stall_lg", "stall_math", "stall_wait", "stall_sleep", "stall_membar"]
cum = 0
for idx, (n, s, r) in enumerate(data):
    cum += n
    if n >= tot * thresh:
        parts = " ".join(f"{nm[6:]}={r[hdr.index(nm)]}" for nm in names if nm in hdr and r[hdr.index(nm)] not in ("0", ""))
        print(f"  {idx:4d} {100*n/tot:5.1f}% (cum {100*cum/tot:5.1f}%) {s[:60]:60s} | {parts}")
