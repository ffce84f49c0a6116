#!/usr/bin/env python
"""tools/profile_step.py --n N --steps K: K ssp-rk3 steps of the bench workload (bench.py's box,
config and field) with nothing else around them -- the command ncu wraps (tools/prof.sh)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--dual", action="store_true", help="one dual (JFNK) evaluation per step instead")
    args = ap.parse_args()
    import bench
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context
    g = bench.box_grid_fast(args.n)
    prim = cases.perturbed_state(g, 3)
    ctx = Context(g, bench.box_config(), device=0)
    tab = cfg.BUTCHER["ssp-rk3"]
    if args.dual:
        fs, dudt = ctx.fs_create(dual=True), ctx.cq_create(dual=True)
        ctx.fs_upload(fs, prim, 1e-3 * prim)
        for _ in range(args.steps):
            ctx.compute_dudt(fs, dudt, True)
    else:
        fs, cq = ctx.fs_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        dt = sys.float_info.max
        for _ in range(args.steps):
            dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
    ctx.sync()
    print("bad cells", ctx.count_bad_cells(fs))
    ctx.close()


if __name__ == "__main__":
    main()
