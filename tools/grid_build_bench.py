"""Time ibis_b200_grid_build on an N^3 jittered hex box given as an SU2 description (vertices, hex
connectivity, six tagged quad lists): one JSON line with seconds and microseconds per cell; the stage
breakdown goes to stderr (IBIS_B200_TIMING=1).  For scale: the reference's GridIO + InterfaceLookup +
init_grid_block path costs ~17 us per cell.

    python tools/grid_build_bench.py --edge 128 [--repeat 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=128)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    os.environ.setdefault("IBIS_B200_TIMING", "1")
    from ibis_b200 import lib, su2
    mesh = su2.box_hex(args.edge, jitter=0.1)
    tags = sorted(mesh.markers)
    best = None
    for _ in range(args.repeat):
        t0 = time.time()
        g = lib.build_grid(mesh, tags)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    n = args.edge ** 3
    print(json.dumps({"tool": "grid_build_bench", "cells": n, "faces": int(g.n_faces), "seconds": best,
                      "us_per_cell": best / n * 1e6, "includes": "ctypes copies of the result arrays",
                      "volume_sum": float(g.cell_volume.sum())}))


if __name__ == "__main__":
    main()
