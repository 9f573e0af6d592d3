#!/usr/bin/env python
"""Opcode histogram and top stall instructions of one kernel from an .ncu-rep captured with --import-source on.
    python tools/ncu_ops.py x.ncu-rep <kernel-name-regex> [n_top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if len(r) > 2 and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
si, ei = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[hi + 1:end]):
    try:
        data.append((int(r[si]), i, r[1].strip(), int(r[ei])))
    except (ValueError, IndexError):
        pass
tot, tex = sum(d[0] for d in data) or 1, sum(d[3] for d in data) or 1
print("%s: %d SASS instructions, %d warp instructions executed, %d stall samples" % (rows[hi - 1][1][:60], len(data), tex, tot))
ops = collections.Counter()
for smp, i, src, ex in data:
    t = src.split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += ex
print("  ".join("%s %.1f%%" % (k, 100 * v / tex) for k, v in ops.most_common(18)))
for smp, i, src, ex in sorted(data, reverse=True)[:ntop]:
    print("%6d %5.1f%%  #%d %s  [%d]" % (smp, 100 * smp / tot, i, src[:90], ex))
