#!/usr/bin/env python
"""One box workload under a profiler: build the N^3 box, run `--steps` RungeKutta steps.
Used by tools/prof.sh under ncu (never for bench numbers)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--dual", action="store_true")
    args = ap.parse_args()
    import bench
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context
    g = bench.box_grid_fast(args.n)
    prim = cases.perturbed_state(g, 3)
    ctx = Context(g, bench.box_config())
    if args.dual:
        fs, dq = ctx.fs_create(True), ctx.cq_create(True)
        ctx.fs_upload(fs, prim, cases.dual_seed(prim))
        for _ in range(args.steps):
            ctx.compute_dudt(fs, dq, True)
        ctx.sync()
    else:
        fs, cq = ctx.fs_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        tab = cfg.BUTCHER["ssp-rk3"]
        dt = sys.float_info.max
        for _ in range(args.steps):
            dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
        ctx.sync()
    ctx.close()


if __name__ == "__main__":
    main()
