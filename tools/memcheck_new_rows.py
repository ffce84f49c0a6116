"""tools/memcheck_new_rows.py: a torch-free walk through the kernels of the SURVEY.md section 8(f) rows, small enough
to run under `compute-sanitizer --tool memcheck` in seconds: native grid build (k_grid_build<0..4>), the
boundary_layer_profile / constant_flux actions (k_bc, k_face_flux, k_integral), the moving-grid evaluation
(k_face_vel, k_face_flux_moving, k_gcl), a vertex move with the geometry recomputed (k_move_vertices,
k_geom_update<0..3>, k_wls_r, k_cell_static), and the reference checks of each."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import parity  # noqa: E402
from ibis_b200 import lib as libmod  # noqa: E402


def hot():
    """--hot: the headline kernels instead (k_bc, k_grad_sides, k_flux_tile, k_rk_update, reductions; dual gather path)"""
    from ibis_b200 import config as cfg
    parity.run_case("box3d_ausmdv")
    mesh, config, g, prim = cases.build_case("box3d_regular")
    ctx = libmod.Context(g, config)
    fs, cq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.prim_to_cons(cq, fs)
    tab = cfg.BUTCHER["ssp-rk3"]
    dt = sys.float_info.max
    for _ in range(2):
        dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
    assert ctx.count_bad_cells(fs) == 0
    ctx.close()
    print("memcheck walk (hot path) ok")


def main():
    if "--hot" in sys.argv:
        return hot()
    for name in ("box3d_profile_cflux", "quad2d_cflux_viscous"):
        parity.run_case(name)
    for name in ("mixed3d_walls", "quad2d_cflux_viscous"):
        mesh, config, g, prim, vvel = cases.build_moving_case(name)
        gn = libmod.build_grid(mesh, list(config["grid"]["boundaries"].keys()))
        gold = dict(np.load(os.path.join(parity.GOLDEN, "moving_" + name + ".npz")))
        ctx = libmod.Context(gn, config)
        ctx.set_face_vertices(gn.face_vertices, len(mesh.vertices))
        ctx.set_cell_vertices(mesh.cell_types, mesh.cell_vertices)
        ctx.set_vertex_positions(mesh.vertices)
        ctx.set_vertex_velocities(vvel)
        fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        ctx.compute_dudt_moving(fs, cq, dq, True)
        d = ctx.cq_download(dq)[:gn.n_cells]
        if config["viscous_flux"]["enabled"]:
            assert np.allclose(d, gold["dudt"], rtol=1e-9, atol=1e-3)
        else:
            assert np.array_equal(d, gold["dudt"]), name
        ctx.move_vertices(1e-6, True)
        ctx.compute_dudt_moving(fs, cq, dq, True)
        ctx.estimate_dt(fs)
        ctx.close()
    print("memcheck walk ok")


if __name__ == "__main__":
    main()
