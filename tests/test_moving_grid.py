"""Moving-grid residual terms (first part): FiniteVolume::compute_dudt(fs, vertex_vel,
cq, grid, dudt, ...) with the rigid_body_translation motion driver -- interface velocities, interface-relative
fluxes and their correction to the inertial frame, constant_flux relative to the interface, the geometric
conservation law -- against the compiled reference (tests/golden/moving_*.npz, tests/golden/make_golden.py
moving): real and dual, second and first order, bit-identical for inviscid configurations.
Each body runs on the host-emulation build here and on the CUDA build under -m gpu."""
import os

import numpy as np
import pytest

import cases
import parity
from emu_util import emu_lib
from ibis_b200 import lib as libmod
from ibis_b200.lib import Context, IbisB200Error

BUILDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _lib(build):
    if build == "emu":
        return emu_lib()
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return None


@pytest.mark.parametrize("build", BUILDS)
@pytest.mark.parametrize("name", sorted(cases.MOVING_CASES))
def test_moving_residual_against_reference(build, name):
    mesh, config, g, prim, vvel = cases.build_moving_case(name)
    gold = dict(np.load(os.path.join(parity.GOLDEN, "moving_" + name + ".npz")))
    assert np.array_equal(prim, gold["prim_in"]) and np.array_equal(vvel, gold["vertex_vel"])
    exact = not config["viscous_flux"]["enabled"]
    S = parity.flux_scale(g, prim, mesh.dim)
    ctx = Context(g, config, lib_path=_lib(build))
    ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
    ctx.set_vertex_velocities(vvel)
    fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
    for allow, sfx in ((True, ""), (False, "_o1")):
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        ctx.compute_dudt_moving(fs, cq, dq, allow)
        d = ctx.cq_download(dq)
        assert not d[g.n_cells:].any()
        parity.assert_close(d[:g.n_cells], gold["dudt" + sfx], S, f"{name}: moving dudt{sfx}", exact)
        parity.assert_close(ctx.fs_download(fs), gold["prim_out" + sfx], 1.0, f"{name}: ghost fill{sfx}", True)
        # GridBlock::face_vel() as the evaluation leaves it: the local-frame round trip of the vertex average
        parity.assert_close(ctx.debug_face_velocities(), gold["face_vel" + sfx], 1.0, f"{name}: face_vel{sfx}", True)
    # it is not the static residual
    ctx.fs_upload(fs, prim)
    ctx.compute_dudt(fs, dq, True)
    assert np.abs(ctx.cq_download(dq)[:g.n_cells] - gold["dudt"]).max() > 1e-3 * np.abs(gold["dudt"]).max()
    # dual evaluation (one Jacobian-vector product direction of the steady solver on a moving grid)
    fsd, cqd, dqd = ctx.fs_create(True), ctx.cq_create(True), ctx.cq_create(True)
    pdu = cases.dual_seed(prim)
    for allow, sfx in ((True, ""), (False, "_o1")):
        ctx.fs_upload(fsd, prim, pdu)
        ctx.prim_to_cons(cqd, fsd)
        ctx.compute_dudt_moving(fsd, cqd, dqd, allow)
        dr, dd = ctx.cq_download(dqd, dual=True)
        parity.assert_close(dr[:g.n_cells], gold["dual_dudt_re" + sfx], S, f"{name}: moving dual re{sfx}", exact)
        parity.assert_close(dd[:g.n_cells], gold["dual_dudt_du" + sfx], 1e-3 * S, f"{name}: moving dual du{sfx}",
                            exact and config["convective_flux"]["flux_calculator"]["type"] != "ldfss2")
    ctx.close()


@pytest.mark.parametrize("build", BUILDS)
def test_zero_velocity_is_the_static_residual(build):
    """vertex velocities 0: interface velocities 0, no GCL contribution -- the moving evaluation equals the static one
    (through the face-flux array instead of the fused tiles) bit for bit"""
    name = "box3d_ausmdv"
    mesh, config, g, prim = cases.build_case(name)
    gold = parity.load_golden(name)
    ctx = Context(g, config, lib_path=_lib(build))
    ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
    ctx.set_vertex_velocities(np.zeros((len(mesh.vertices), 3)))
    fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.prim_to_cons(cq, fs)
    ctx.compute_dudt_moving(fs, cq, dq, True)
    assert np.array_equal(ctx.cq_download(dq)[:g.n_cells], gold["dudt"])
    ctx.close()


def test_moving_api_errors():
    L = emu_lib()
    mesh, config, g, prim = cases.build_case("quad2d_hanel")
    ctx = Context(g, config, lib_path=L)
    fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
    with pytest.raises(IbisB200Error, match="set_face_vertices has not been called"):
        ctx.set_vertex_velocities(np.zeros((len(mesh.vertices), 3)))
    with pytest.raises(IbisB200Error, match="not moving"):
        ctx.compute_dudt_moving(fs, cq, dq)
    bad = np.array(g.face_vertices)
    bad[3, 1] = len(mesh.vertices) + 7
    with pytest.raises(IbisB200Error, match="outside the vertex list"):
        ctx.set_face_vertices(bad, len(mesh.vertices))
    ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
    with pytest.raises(IbisB200Error, match="already set"):
        ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
    ctx.close()


@pytest.mark.parametrize("build", BUILDS)
def test_geometry_after_a_vertex_move(build):
    """move_vertices + the recomputed geometry (GridBlock::compute_geometric_data, grid.h:169-178): after a rigid
    displacement the moving evaluation with the vertices at rest equals the static evaluation of a context built
    on the displaced mesh -- to round-off only, because the reference (and this) keeps the ORIGINAL interface
    orientations while a fresh build recomputes them from the displaced vertices."""
    from ibis_b200 import grid as gridmod
    name = "mixed3d_walls"
    mesh, config, g, prim = cases.build_case(name)
    shift = np.array([0.37, -0.21, 0.11])
    ctx = Context(g, config, lib_path=_lib(build))
    ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
    ctx.set_cell_vertices(mesh.cell_types, mesh.cell_vertices)
    ctx.set_vertex_positions(mesh.vertices)
    ctx.set_vertex_velocities(np.tile(shift, (len(mesh.vertices), 1)))
    ctx.move_vertices(0.25, recompute_geometry=False)
    ctx.move_vertices(0.75)                                   # two moves, one recomputation
    assert np.array_equal(ctx.get_vertex_positions(), (mesh.vertices + shift * 0.25) + shift * 0.75)
    ctx.set_vertex_velocities(np.zeros((len(mesh.vertices), 3)))
    fs, cq, dq = ctx.fs_create(), ctx.cq_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.prim_to_cons(cq, fs)
    ctx.compute_dudt_moving(fs, cq, dq, True)
    moved = ctx.cq_download(dq)[:g.n_cells]
    with pytest.raises(IbisB200Error, match="vertices have moved"):
        ctx.compute_dudt(fs, dq, True)
    ctx.close()
    mesh.vertices = (mesh.vertices + shift * 0.25) + shift * 0.75
    g2 = gridmod.build_grid(mesh, list(config["grid"]["boundaries"].keys()))
    ctx2 = Context(g2, config, lib_path=_lib(build))
    fs, dq = ctx2.fs_create(), ctx2.cq_create()
    ctx2.fs_upload(fs, prim)
    ctx2.compute_dudt(fs, dq, True)
    fresh = ctx2.cq_download(dq)[:g.n_cells]
    ctx2.close()
    S = parity.flux_scale(g, prim, mesh.dim)
    assert (np.abs(moved - fresh) <= 1e-9 * np.maximum(np.abs(fresh), S)).all()
