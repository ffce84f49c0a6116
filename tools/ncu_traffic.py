#!/usr/bin/env python
"""tools/ncu_traffic.py RAW.csv N PATH [OUT]: dram__bytes_read.sum + dram__bytes_write.sum per launch of pass G' and
pass F' from the raw page of an `ncu --set full` capture of tools/profile_step.py --n N (tools/prof.sh), per cell,
into profiles/traffic.json -- what bench.py reports as roofline.traffic (scaled by its cell count).
PATH = sides | gather."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    raw, n, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "traffic.json")
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    name = hdr.index("Kernel Name")
    res = {}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = "G" if "grad" in r[name] else "F" if ("flux" in r[name] or "residual" in r[name]) else None
        if k is None or k in res:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * UNIT[units[i]]
        res[k] = tot / n ** 3
        res[k + "_kernel"] = r[name][:80]
    res["source"] = os.path.relpath(raw, ROOT) + f" (n = {n}, {n ** 3} cells)"
    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    data[path] = res
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
