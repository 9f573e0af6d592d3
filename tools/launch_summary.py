#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), len(hdr) - 1
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
    except Exception:
        pass
for k, v in agg.items():
    print("%-72s n=%4d  sum=%12.1f  mean=%10.1f  min=%10.1f" % (k, len(v), sum(v), sum(v) / len(v), min(v)))
