"""Whole-solver parity: the C++ host solvers (ibis_b200/host) above the C ABI against the final
flow snapshots of the real reference (tests/golden/solve_*.npz, made by make_solve_golden.py).

Tolerance: north_star -- flow fields after N steps within 1e-9 (relative to the field's scale).
Explicit inviscid runs are in fact bit-identical (same IEEE operations in the same order, dt from
an exact min-reduction); viscous runs differ through pow() in Sutherland's law, steady runs through
the summation order of dot products and norms."""
import os

import numpy as np

import solve_cases as sc
from ibis_b200 import host
from ibis_b200.lib import Context

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9


def run_solve_case(name, lib_path=None, host_lib=None, fused_step=True, exact=None, rtol=RTOL):
    mesh, config, g, prim0, snap = sc.build_solve_case(name)
    gold = dict(np.load(os.path.join(GOLDEN, "solve_" + name + ".npz")))
    ctx = Context(g, config, lib_path=lib_path)
    motion = config["grid"].get("motion", {})
    if motion.get("enabled"):
        # what RungeKutta's constructor wires for a moving grid (runge_kutta.cpp:58-70); the rigid-body driver
        # hands every vertex the same velocity (rigid_body_translation.cpp:14-27)
        v = motion["velocity"]
        ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
        ctx.set_cell_vertices(mesh.cell_types, mesh.cell_vertices)
        ctx.set_vertex_positions(mesh.vertices)
        ctx.set_vertex_velocities(np.tile([v["x"], v["y"], v["z"]], (len(mesh.vertices), 1)))
    try:
        out, rep = host.solve(ctx, config, prim0, fused_step=fused_step, host_lib=host_lib)
    finally:
        ctx.close()
    assert rep["status"] == int(gold["rc"]) == 0
    n = g.n_cells
    cols = {"T": 2, "p": 1, "vx": 4, "vy": 5, "vz": 6}
    vscale = np.sqrt(out[:n, 4] ** 2 + out[:n, 5] ** 2 + out[:n, 6] ** 2).max()
    worst = 0.0
    for k, col in cols.items():
        if "final_" + k not in gold:
            continue
        a, b = out[:n, col], gold["final_" + k]
        scale = vscale if k.startswith("v") else np.abs(b).max()
        err = np.abs(a - b).max() / scale
        worst = max(worst, err)
        assert err <= rtol, f"{name}: {k} differs by {err:.3e} relative after {rep['steps']} steps"
        if exact:
            assert np.array_equal(a, b), f"{name}: {k} not bit-identical (max rel {err:.3e})"
    if config["solver"]["name"] == "runge_kutta":
        assert abs(rep["t"] - float(gold["final_time"])) <= 1e-12 * abs(float(gold["final_time"]))
    else:
        # the reference plots with t = step index of the last step (steady_state.cpp:264-270)
        assert rep["steps"] - 1 == int(float(gold["final_time"])), (rep["steps"], float(gold["final_time"]))
        if "log_gmres_diagnostics" in gold:
            pass
    return worst, rep
