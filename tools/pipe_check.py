#!/usr/bin/env python
"""tools/pipe_check.py --n N [--steps K]: the pipelined sides path against the gather path on an N^3 perturbed
box (bit for bit), residual and K ssp-rk3 steps; progress on stderr so that a hang can be located."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def say(*a):
    print(*a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import bench
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context
    g = bench.box_grid_fast(args.n)
    prim = cases.perturbed_state(g, 3)
    tab = cfg.BUTCHER["ssp-rk3"]
    out = {}
    for sides in ("1", "0"):
        os.environ["IBIS_B200_SIDES"] = sides
        say(f"n={args.n} sides={sides}: create")
        ctx = Context(g, bench.box_config(), device=0)
        fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        say("  compute_dudt")
        ctx.compute_dudt(fs, dq, True)
        ctx.sync()
        r = ctx.cq_download(dq)
        say("  second compute_dudt")
        ctx.compute_dudt(fs, dq, True)
        ctx.sync()
        r2 = ctx.cq_download(dq)
        assert np.array_equal(r, r2, equal_nan=True), "evaluation not repeatable"
        ctx.prim_to_cons(cq, fs)
        dt = sys.float_info.max
        for k in range(args.steps):
            say(f"  rk_step {k}")
            dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
        ctx.sync()
        out[sides] = (r, ctx.fs_download(fs), dt)
        ctx.close()
    same_r = np.array_equal(out["1"][0], out["0"][0])
    same_f = np.array_equal(out["1"][1], out["0"][1])
    print(f"n={args.n} residual equal: {same_r}  state after {args.steps} steps equal: {same_f}  dt {out['1'][2]} {out['0'][2]}")
    if not (same_r and same_f and out["1"][2] == out["0"][2]):
        d = np.argwhere(out["1"][0] != out["0"][0])
        print("first differing entries", d[:10].tolist(), "of", len(d))
        sys.exit(1)


if __name__ == "__main__":
    main()
