#!/usr/bin/env python
"""Opcode mix of one kernel from `ncu --page source --csv` output: executed warp
instructions and stall samples per SASS opcode."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
si, ei, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, samp = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip())
    op = s.split()[0].split(".")[0]
    agg[op] += int(r[ei])
    samp[op] += int(r[ss])
tot, ts = sum(agg.values()), max(sum(samp.values()), 1)
print("total warp-inst", tot, " per warp", tot / nwarps)
for op, n in agg.most_common(32):
    print(f"{op:12s} {n:12d} {n / tot * 100:5.1f}%  per-warp {n / nwarps:8.1f}  samples {samp[op] / ts * 100:5.1f}%")
