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
    # round 2: asynchronous steps, the snapshot ring (two copy streams), a dual evaluation through the pipelined
    # kernels, one GMRES iteration
    host = [ctx.pinned_array((5, ctx.n_cells)) for _ in range(3)]
    for q, a in enumerate(ctx.flow_write_native(fs)):
        host[0][q] = a
    ctx.snapshot_upload_begin(0, host[0])
    for k in range(4):
        ctx.snapshot_upload_end(k & 1, fs)
        if k < 3:
            ctx.snapshot_upload_begin((k & 1) ^ 1, host[0])
        ctx.prim_to_cons(cq, fs)
        ctx.rk_step_async(fs, cq, tab["a"], tab["b"], 0.5, 1.0)
        if k >= 2:
            ctx.snapshot_download_end(k & 1)
        ctx.snapshot_download_begin(k & 1, fs, host[1 + (k & 1)])
    ctx.snapshot_download_end(0)
    ctx.snapshot_download_end(1)
    assert np.isfinite(host[1]).all() and np.isfinite(host[2]).all() and ctx.rk_last_dt()[0] > 0
    n = ctx.n_cells
    fd, qd, td, dd = ctx.fs_create(dual=True), ctx.cq_create(dual=True), ctx.cq_create(dual=True), ctx.cq_create(dual=True)
    ctx.fs_upload(fd, prim, 1e-3 * prim)
    ctx.prim_to_cons(qd, fd)
    v, w, basis = ctx.vec_create(n * 5), ctx.vec_create(n * 5), ctx.basis_create(n * 5, 3)
    ctx.vec_upload(v, np.full(n * 5, 1.0 / np.sqrt(n * 5.0)))
    ctx.basis_set_col(basis, 0, v)
    for j in range(2):
        ctx.jfnk_set_dual(qd, v, td)
        ctx.cons_to_prim(td, fd)
        ctx.compute_dudt(fd, dd, True)
        ctx.jfnk_set_vector(dd, v, 1e-4, w)
        assert np.isfinite(ctx.arnoldi_mgs(basis, j, w, v)).all()
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
