"""The known-answer tests the REFERENCE's own unit-test suite holds for this path,
replayed through the C ABI -- on the host-emulation build here, on the CUDA build under -m gpu:

  grid/grid.cpp:184-318       cell faces, outsigns, neighbours, ghost-cell centres of the 3 x 3 quad grid
  grid/cell.cpp:68-96         cell volumes and centres of the same grid
  grid/grid.cpp:145-182       the 27-hex cube: volumes 1/27, unit tangents and normals
  grid/gradient.cpp:29-82     weighted-least-squares gradient of a linear field: (1.0, 0.5)
  linear_algebra/dense_linear_algebra.cpp:27-196   norm2, scale, add_scaled_vector, gemv, dot
  (GMRES / FGMRES 5 x 5 systems and LinearSchedule: tests/test_emu_solvers.py, test_gpu_solvers.py;
   CubicSpline: tests/test_boundary_actions.py)

The reference's two fixture grids (src/grid/test/grid.su2, cube.su2) are regenerated here as text -- they
are the canonical 3 x 3 and 3 x 3 x 3 lattices -- and, when /root/reference is present (build container
only), read back with the native SU2 reader next to the reference's own files to show they hold the same
mesh, quirks included (comment and blank lines, `NELEM` before `NPOIN`, trailing element indices)."""
import os

import numpy as np
import pytest

from emu_util import emu_lib
from ibis_b200 import config as cfg
from ibis_b200 import lib as libmod
from ibis_b200.lib import Context

BUILDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
REF_TEST_DIR = "/root/reference/src/grid/test"


def _lib(build):
    if build == "emu":
        return emu_lib()
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return None


def _grid_su2_text():
    """the 3 x 3 unit-square lattice of src/grid/test/grid.su2, its marker order and its decorations"""
    out = ["% test comment followed by a blank line, to check if comments and blank lines work.", "", "NDIME= 2", "NPOIN= 16"]
    out += [f"{float(i)} {float(j)}" for j in range(4) for i in range(4)]
    out.append("NELEM= 9")
    out += [f"9 {4 * j + i} {4 * j + i + 1} {4 * j + i + 5} {4 * j + i + 4}" for j in range(3) for i in range(3)]
    out.append("NMARK= 4")
    marks = [("slip_wall_bottom", [(i, i + 1) for i in range(3)]), ("outflow", [(4 * j + 3, 4 * j + 7) for j in range(3)]),
             ("slip_wall_top", [(12 + i, 13 + i) for i in range(3)]), ("inflow", [(4 * j, 4 * j + 4) for j in range(3)])]
    for tag, elems in marks:
        out += [f"MARKER_TAG={tag}", "MARKER_ELEMS= 3"] + [f"3 {a} {b}" for a, b in elems]
    out.append("Footer")
    return "\n".join(out) + "\n"


def _cube_su2_text():
    """the 3 x 3 x 3 unit-cube lattice of src/grid/test/cube.su2: elements first, trailing indices"""
    c = np.linspace(0.0, 1.0, 4)
    vid = lambda i, j, k: i + 4 * j + 16 * k                                      # noqa: E731
    out = ["NDIME= 3", "", "NELEM= 27"]
    n = 0
    for k in range(3):
        for j in range(3):
            for i in range(3):
                v = [vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
                     vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)]
                out.append("12 " + " ".join(map(str, v)) + f" {n}")
                n += 1
    out += ["", "NPOIN= 64"]
    out += [f"{c[i]:.18e} {c[j]:.18e} {c[k]:.18e} {vid(i, j, k)}" for k in range(4) for j in range(4) for i in range(4)]
    quads = {"west": [(vid(0, j, k), vid(0, j, k + 1), vid(0, j + 1, k + 1), vid(0, j + 1, k)) for k in range(3) for j in range(3)],
             "east": [(vid(3, j, k), vid(3, j + 1, k), vid(3, j + 1, k + 1), vid(3, j, k + 1)) for k in range(3) for j in range(3)],
             "south": [(vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1)) for k in range(3) for i in range(3)],
             "north": [(vid(i, 3, k), vid(i, 3, k + 1), vid(i + 1, 3, k + 1), vid(i + 1, 3, k)) for k in range(3) for i in range(3)],
             "bottom": [(vid(i, j, 0), vid(i, j + 1, 0), vid(i + 1, j + 1, 0), vid(i + 1, j, 0)) for j in range(3) for i in range(3)],
             "top": [(vid(i, j, 3), vid(i + 1, j, 3), vid(i + 1, j + 1, 3), vid(i, j + 1, 3)) for j in range(3) for i in range(3)]}
    out.append("NMARK= 6")
    for tag, elems in quads.items():
        out += [f"MARKER_TAG= {tag}", "MARKER_ELEMS= 9"] + ["9 " + " ".join(map(str, q)) for q in elems]
    return "\n".join(out) + "\n"


def _read(tmp_path, text, build):
    p = str(tmp_path / "g.su2")
    with open(p, "w") as f:
        f.write(text)
    return libmod.read_su2(p, lib_path=_lib(build))


@pytest.mark.skipif(not os.path.isdir(REF_TEST_DIR), reason="the reference tree is only in the build container")
def test_regenerated_fixture_grids_are_the_reference_files(tmp_path):
    for text, name in ((_grid_su2_text(), "grid.su2"), (_cube_su2_text(), "cube.su2")):
        mine = _read(tmp_path, text, "emu")
        ref = libmod.read_su2(os.path.join(REF_TEST_DIR, name), lib_path=emu_lib())
        # cube.su2 was written by a generator whose thirds differ from linspace's in the last ulp
        assert mine.dim == ref.dim and np.abs(mine.vertices - ref.vertices).max() < 1e-15
        assert np.array_equal(mine.cell_types, ref.cell_types) and np.array_equal(mine.cell_vertices, ref.cell_vertices)
        assert list(mine.markers) == list(ref.markers)
        for t in ref.markers:
            a = np.sort(np.sort(mine.markers[t][1], axis=1), axis=0)       # same boundary faces (vertex order within
            b = np.sort(np.sort(ref.markers[t][1], axis=1), axis=0)        # a face and face order within a tag are free)
            assert np.array_equal(a, b), (name, t)


@pytest.mark.parametrize("build", BUILDS)
def test_grid_kats_3x3(build, tmp_path):
    mesh = _read(tmp_path, _grid_su2_text(), build)
    g = libmod.build_grid(mesh, None, lib_path=_lib(build))
    assert g.n_cells == 9 and g.n_faces == 24 and g.n_ghost == 12
    off = g.cell_face_offsets
    faces = [list(g.cell_face_ids[off[i]:off[i + 1]]) for i in range(9)]
    assert faces == [[0, 1, 2, 3], [4, 5, 6, 1], [7, 8, 9, 5], [2, 10, 11, 12], [6, 13, 14, 10], [9, 15, 16, 13],
                     [11, 17, 18, 19], [14, 20, 21, 17], [16, 22, 23, 20]]                      # grid.cpp:199-215
    outs = [list(g.cell_outsigns[off[i]:off[i + 1]]) for i in range(9)]
    assert outs == [[1, 1, 1, 1], [1, 1, 1, -1], [1, 1, 1, -1], [-1, 1, 1, 1], [-1, 1, 1, -1], [-1, 1, 1, -1],
                    [-1, 1, 1, 1], [-1, 1, 1, -1], [-1, 1, 1, -1]]                              # grid.cpp:217-232
    nb = lambda c, j: int(g.cell_neighbours[off[c] + j])                                        # noqa: E731
    assert (nb(0, 1), nb(0, 2), nb(4, 0), nb(8, 3)) == (1, 3, 1, 7)                             # grid.cpp:234-244
    cen = g.cell_centroid
    want = {"inflow": [(-0.5, 0.5), (-0.5, 1.5), (-0.5, 2.5)], "slip_wall_bottom": [(0.5, -0.5), (1.5, -0.5), (2.5, -0.5)],
            "outflow": [(3.5, 0.5), (3.5, 1.5), (3.5, 2.5)], "slip_wall_top": [(0.5, 3.5), (1.5, 3.5), (2.5, 3.5)]}
    for tag, pts in want.items():                                                               # grid.cpp:246-318
        for k, (x, y) in enumerate(pts):
            gc = g.tags[tag]["ghosts"][k]
            assert tuple(cen[gc]) == (x, y, 0.0), (tag, k)
    assert np.abs(g.cell_volume - 1.0).max() < 1e-14                                            # cell.cpp:68-80
    xs, ys = [0.5, 1.5, 2.5] * 3, [0.5] * 3 + [1.5] * 3 + [2.5] * 3                             # cell.cpp:82-96
    assert np.abs(cen[:9, 0] - xs).max() < 1e-14 and np.abs(cen[:9, 1] - ys).max() < 1e-14


@pytest.mark.parametrize("build", BUILDS)
def test_grid_kats_cube(build, tmp_path):
    mesh = _read(tmp_path, _cube_su2_text(), build)
    g = libmod.build_grid(mesh, None, lib_path=_lib(build))
    assert g.n_cells == 27                                                                      # grid.cpp:145-155
    np.testing.assert_allclose(g.cell_volume, 1.0 / 27, rtol=1e-12)
    for k in ("face_tan1", "face_tan2", "face_norm"):                                           # grid.cpp:157-182
        np.testing.assert_allclose(np.sqrt((getattr(g, k) ** 2).sum(axis=1)), 1.0, rtol=1e-12)


@pytest.mark.parametrize("build", BUILDS)
def test_wls_gradient_of_a_linear_field(build, tmp_path):
    """gradient.cpp:29-82: values 1, 2, 3 / 1.5, 2.5, 3.5 / 2, 3, 4 in the cells and the matching ghost values are
    x + y/2 + 1/4 at the centroids; every cell's gradient is (1.0, 0.5).  The field rides in as the density of
    a state whose boundaries carry no ghost-cell action, so the uploaded ghost values stay."""
    mesh = _read(tmp_path, _grid_su2_text(), build)
    tags = sorted(mesh.markers)
    g = libmod.build_grid(mesh, tags, lib_path=_lib(build))
    values = g.cell_centroid[:, 0] + 0.5 * g.cell_centroid[:, 1] + 0.25
    assert list(values[:9]) == [1.0, 2.0, 3.0, 1.5, 2.5, 3.5, 2.0, 3.0, 4.0]
    assert list(values[g.tags["slip_wall_bottom"]["ghosts"]]) == [0.5, 1.5, 2.5]
    assert list(values[g.tags["slip_wall_top"]["ghosts"]]) == [2.5, 3.5, 4.5]
    assert list(values[g.tags["inflow"]["ghosts"]]) == [0.0, 0.5, 1.0]
    assert list(values[g.tags["outflow"]["ghosts"]]) == [4.0, 4.5, 5.0]
    no_action = {"pre_reconstruction": [], "post_convective_flux": [], "pre_viscous_grad": [], "ghost_cells": True}
    config = cfg.make_config(2, {t: dict(no_action) for t in tags}, flux_calculator="hanel", limiter="unlimited")
    n = g.n_cells + g.n_ghost
    prim = np.zeros((n, 7))
    prim[:, 0] = values + 1.0                      # a density must be positive: the constant does not change the gradient
    prim[:, 2] = 300.0
    prim[:, 1] = prim[:, 0] * 287.0 * 300.0
    prim[:, 3] = 717.5 * 300.0
    prim[:, 4] = 100.0
    ctx = Context(g, config, lib_path=_lib(build))
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.compute_dudt(fs, dq, True)
    grad, _ = ctx.debug_gradients(0)
    ctx.close()
    np.testing.assert_allclose(grad[:, 0], 1.0, rtol=1e-12)
    np.testing.assert_allclose(grad[:, 1], 0.5, rtol=1e-12)
    assert np.abs(grad[:, 2]).max() < 1e-12


@pytest.mark.parametrize("build", BUILDS)
def test_dense_linear_algebra_kats(build, tmp_path):
    mesh = _read(tmp_path, _grid_su2_text(), build)
    tags = sorted(mesh.markers)
    g = libmod.build_grid(mesh, tags, lib_path=_lib(build))
    ctx = Context(g, cfg.make_config(2, {t: cfg.supersonic_outflow() for t in tags}), lib_path=_lib(build))
    small = np.array([1.0, 2.0, -3.0, 1.5])                                  # small_test_vector
    x, y = ctx.vec_create(4), ctx.vec_create(4)
    ctx.vec_upload(x, small)
    assert ctx.vec_norm2(x) == pytest.approx(np.sqrt(16.25), rel=1e-15)      # Ibis::norm2
    ctx.vec_scale(x, y, 2.0)                                                 # Ibis::scale
    assert list(ctx.vec_download(y, 4)) == [2.0, 4.0, -6.0, 3.0]
    ctx.vec_scale(x, x, 2.0)                                                 # Ibis::scale_in_place
    assert list(ctx.vec_download(x, 4)) == [2.0, 4.0, -6.0, 3.0]
    ctx.vec_upload(x, small)
    ctx.vec_upload(y, small)
    ctx.vec_axpy(x, y, 2.0)                                                  # Ibis::add_scaled_vector
    assert list(ctx.vec_download(x, 4)) == [3.0, 6.0, -9.0, 4.5]
    a, b = ctx.vec_create(3), ctx.vec_create(3)
    ctx.vec_upload(a, [1.0, 2.0, 3.0])
    ctx.vec_upload(b, [2.0, 3.0, 4.0])
    assert ctx.vec_dot(a, b) == 20.0                                         # Ibis::dot
    basis, col, res = ctx.basis_create(3, 5), ctx.vec_create(3), ctx.vec_create(3)
    for k in range(5):                                                       # Ibis::gemv: lhs(i, k) = 5 i + k + 1
        ctx.vec_upload(col, [k + 1.0, k + 6.0, k + 11.0])
        ctx.basis_set_col(basis, k, col)
    ctx.basis_gemv(basis, 5, [1.0, 2.0, 3.0, 4.0, 5.0], res)
    assert list(ctx.vec_download(res, 3)) == [55.0, 130.0, 205.0]
    ctx.close()


@pytest.mark.parametrize("build", BUILDS)
def test_dual_number_kats(build, tmp_path):
    """util/dual.cpp:4-222 through the device's own dual operators (ibis_b200_debug_dual_ops)"""
    mesh = _read(tmp_path, _grid_su2_text(), build)
    tags = sorted(mesh.markers)
    g = libmod.build_grid(mesh, tags, lib_path=_lib(build))
    ctx = Context(g, cfg.make_config(2, {t: cfg.supersonic_outflow() for t in tags}), lib_path=_lib(build))
    names = ["x+y", "x+=y", "x+r", "x-y", "x-=y", "x-r", "r-x", "x*y", "x*=y", "x*r", "x*=r", "r*x", "x/y", "x/=y",
             "x/r", "x/=r", "r/x", "sqrt", "abs", "abs(x*x)", "x*x+2x", "x*x-sqrt(x)"]

    def ops(x, y=(2.0, 3.0), r=2.0):
        return dict(zip(names, map(tuple, ctx.debug_dual_ops(x, y, r))))

    o = ops((3.0, 1.0))
    want = {"x+y": (5.0, 4.0), "x+=y": (5.0, 4.0), "x+r": (5.0, 1.0), "x-y": (1.0, -2.0), "x-=y": (1.0, -2.0),
            "x-r": (1.0, 1.0), "r-x": (-1.0, -1.0), "x*y": (6.0, 11.0), "x*=y": (6.0, 11.0), "x*r": (6.0, 2.0),
            "x*=r": (6.0, 2.0), "r*x": (6.0, 2.0), "x/y": (1.5, -1.75), "x/=y": (1.5, -1.75), "x/r": (1.5, 0.5),
            "x/=r": (1.5, 0.5)}
    for k, v in want.items():
        assert o[k] == pytest.approx(v, rel=1e-15), k
    assert ops((4.0, 1.0))["r/x"] == pytest.approx((0.5, -0.125), rel=1e-15)
    assert ops((9.0, 3.0))["sqrt"] == pytest.approx((3.0, 0.5), rel=1e-15)
    assert ops((2.0, 1.0))["x*x+2x"] == pytest.approx((8.0, 6.0), rel=1e-15)
    assert ops((9.0, 1.0))["x*x-sqrt(x)"] == pytest.approx((78.0, 18.0 - 1.0 / 6.0), rel=1e-15)
    assert ops((-1.0, 1.0))["abs"] == (1.0, -1.0) and ops((1.0, 1.0))["abs"] == (1.0, 1.0)
    assert ops((-1.0, 1.0))["abs(x*x)"] == (1.0, -2.0) and ops((1.0, 1.0))["abs(x*x)"] == (1.0, 2.0)
    ctx.close()
