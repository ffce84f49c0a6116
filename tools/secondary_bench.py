#!/usr/bin/env python
"""tools/secondary_bench.py [--n 160]: residual-evaluation throughput of the variants of the path that the
headline bench does not time (SURVEY.md section 8d "variants"): first order (the JFNK preconditioner
residual), dual numbers (the exact Jacobian-vector product of the steady solver), the viscous path and
the other flux calculators, on the n^3 box.  One JSON line per variant: evals/s and the fraction of the
measured HBM roofline on that variant's algorithmic bytes."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=160)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import bench
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context
    peak, _ = bench.measured_peak()
    g = bench.box_grid_fast(args.n)
    prim = cases.perturbed_state(g, 3)
    n = g.n_cells
    variants = [
        ("real order 2 ausmdv (headline)", dict(flux_calculator="ausmdv"), False, True, 1030),
        ("real order 1 (allow_reconstruction = false)", dict(flux_calculator="ausmdv"), False, False, 382),
        ("dual order 2 (J.v of the steady solver)", dict(flux_calculator="ausmdv"), True, True, 1470),
        ("dual order 1 (preconditioner J.v)", dict(flux_calculator="ausmdv"), True, False, 462),
        ("real order 2 hanel", dict(flux_calculator="hanel"), False, True, 1030),
        ("real order 2 ldfss2", dict(flux_calculator="ldfss2"), False, True, 1030),
        ("real order 2 rusanov", dict(flux_calculator="rusanov"), False, True, 1030),
        ("real order 2 viscous (no-slip walls)", dict(flux_calculator="ausmdv", viscous=True), False, True, 1360),
    ]
    for name, kw, dual, allow, alg in variants:
        bcs = cases.box3_visc_bcs() if kw.get("viscous") else cases.box3_bcs()
        config = cfg.make_config(3, bcs, reconstruction_order=2, solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5), **kw)
        ctx = Context(g, config, device=0)
        fs, dq = ctx.fs_create(dual=dual), ctx.cq_create(dual=dual)
        ctx.fs_upload(fs, prim, 1e-3 * prim if dual else None)
        for _ in range(3):
            ctx.compute_dudt(fs, dq, allow)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            ctx.compute_dudt(fs, dq, allow)
        ctx.sync()
        t = (time.perf_counter() - t0) / args.reps
        print(json.dumps({"variant": name, "cells": int(n), "ms_per_evaluation": t * 1e3, "evals_per_s": n / t,
                          "algorithmic_bytes_per_cell": alg, "frac_of_measured_hbm": alg * n / t / 1e9 / peak}))
        ctx.close()


if __name__ == "__main__":
    main()
