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
    ap.add_argument("--n-moving", type=int, default=96)
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

    # constant_flux inflow: the evaluation goes through the stored face-flux array (k_face_flux -> k_bc -> k_integral)
    bcs = cases.box3_bcs()
    bcs["west"] = cfg.constant_flux(cases.FS3)
    config = cfg.make_config(3, bcs, flux_calculator="ausmdv")
    ctx = Context(g, config, device=0)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    for _ in range(3):
        ctx.compute_dudt(fs, dq, True)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.compute_dudt(fs, dq, True)
    ctx.sync()
    t = (time.perf_counter() - t0) / args.reps
    print(json.dumps({"variant": "real order 2 ausmdv, constant_flux inflow (face-flux-array path)", "cells": int(n),
                      "ms_per_evaluation": t * 1e3, "evals_per_s": n / t, "algorithmic_bytes_per_cell": 1030,
                      "frac_of_measured_hbm": 1030 * n / t / 1e9 / peak}))
    ctx.close()

    # moving grid (rigid-body translation) on a natively built grid: interface velocities, relative fluxes, GCL
    from ibis_b200 import lib as libmod
    from ibis_b200 import su2
    m = args.n_moving
    mesh = su2.box_hex(m, jitter=0.1)
    t0 = time.perf_counter()
    gm = libmod.build_grid(mesh, sorted(mesh.markers))
    t_build = time.perf_counter() - t0
    config = cfg.make_config(3, cases.box3_bcs(), flux_calculator="ausmdv", motion=cfg.rigid_body_translation(120.0, -75.0, 40.0))
    primm = cases.perturbed_state(gm, 3)
    ctx = Context(gm, config, device=0)
    ctx.set_face_vertices(gm.face_vertices, len(mesh.vertices))
    ctx.set_cell_vertices(mesh.cell_types, mesh.cell_vertices)
    ctx.set_vertex_positions(mesh.vertices)
    import numpy as np
    ctx.set_vertex_velocities(np.tile([120.0, -75.0, 40.0], (len(mesh.vertices), 1)))
    fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
    ctx.fs_upload(fs, primm)
    ctx.prim_to_cons(cq, fs)
    for _ in range(3):
        ctx.compute_dudt_moving(fs, cq, dq, True)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.compute_dudt_moving(fs, cq, dq, True)
    ctx.sync()
    t = (time.perf_counter() - t0) / args.reps
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.move_vertices(1e-9, True)
    ctx.sync()
    t_move = (time.perf_counter() - t0) / args.reps
    nm = gm.n_cells
    print(json.dumps({"variant": "moving grid, real order 2 ausmdv (compute_dudt_moving: face velocities, relative fluxes, GCL)",
                      "cells": int(nm), "ms_per_evaluation": t * 1e3, "evals_per_s": nm / t,
                      "ms_per_vertex_move_and_geometry": t_move * 1e3, "native_grid_build_s": t_build}))
    ctx.close()


if __name__ == "__main__":
    main()
