#!/usr/bin/env python
"""Per-launch rows (kernel, grid, block, duration in us) of an ncu --csv launch list.
    python tools/launch_rows.py gpurun_out/x.csv [substring ...]"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
pats = sys.argv[2:]
for row in csv.DictReader(lines):
    n = row["Kernel Name"]
    if pats and not any(p in n for p in pats):
        continue
    print("%-34s %-16s %-12s %10.1f" % (n[:34], row["Grid Size"], row["Block Size"], float(row["Metric Value"]) / 1e3))
