#!/usr/bin/env python
"""tools/ncu_hot.py SOURCE.csv [N]: the N SASS instructions with the most stall samples of an
`ncu --page source --csv` export, plus totals (samples, executed warp instructions)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ilsb = hdr.index("stall_long_sb")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # further launches of the capture: the first one is enough
    if len(r) == len(hdr):
        data.append(r)
tot = sum(int(r[isamp]) for r in data)
texe = sum(int(r[iex]) for r in data)
print(f"{len(data)} SASS instructions, {tot} samples, {texe} warp instructions executed")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} {100 * int(r[isamp]) / tot:5.1f}% long_sb {r[ilsb]:>6s} exec {r[iex]:>8s}  {r[isrc].strip()[:100]}")
