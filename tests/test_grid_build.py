"""ibis_b200_grid_build: SU2 mesh -> the arrays of the reference GridBlock by
one sort + geometry kernels, against (i) the arrays the REAL reference built for the same meshes
(tests/golden/*.npz, grid_* entries, written by tests/golden/make_golden.py from GridBlock<real>) and
(ii) the numpy mirror ibis_b200/grid.py for the arrays the fixtures do not hold -- bit for bit, every
mesh kind of tests/cases.py (hex, jittered hex, wedge/pyramid/tet mixes, quads, triangles, O-mesh).
Each body runs on the host-emulation build here and on the CUDA build under -m gpu."""
import numpy as np
import pytest

import cases
import parity
from emu_util import emu_lib
from ibis_b200 import grid as gridmod
from ibis_b200 import lib as libmod
from ibis_b200 import su2
from ibis_b200.lib import Context, IbisB200Error

BUILDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
MESH_CASES = ["box3d_ausmdv", "box3d_regular", "box3d_viscous", "mixed3d_walls", "tet3d", "tri2d", "quad2d_ausmdv",
              "wedge2d", "naca_omesh"]
ARRAYS = ("face_left", "face_right", "face_area", "face_norm", "face_tan1", "face_tan2", "face_centre",
          "cell_face_offsets", "cell_face_ids", "cell_outsigns", "cell_neighbours", "cell_volume", "cell_centroid",
          "face_vertices")


def _lib(build):
    if build == "emu":
        return emu_lib()
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return None


def _mesh_of(name):
    mesh_f, bcs_f, _, _ = cases.CASES[name]
    return mesh_f(), list(bcs_f().keys())


@pytest.mark.parametrize("build", BUILDS)
@pytest.mark.parametrize("name", MESH_CASES)
def test_native_grid_equals_reference_gridblock(build, name):
    mesh, tags = _mesh_of(name)
    g = libmod.build_grid(mesh, tags, lib_path=_lib(build))
    gold = parity.load_golden(name)
    for k in ("face_left", "face_right", "face_area", "face_norm", "cell_volume", "cell_centroid",
              "cell_face_ids", "cell_outsigns", "cell_neighbours"):
        assert np.array_equal(np.asarray(getattr(g, k)), gold["grid_" + k]), f"{name}: {k} differs from the reference's"
    m = gridmod.build_grid(mesh, tags)
    for k in ARRAYS:
        assert np.array_equal(np.asarray(getattr(g, k)), np.asarray(getattr(m, k))), f"{name}: {k} differs from the mirror"
    assert sorted(g.tags) == sorted(m.tags)
    for t in m.tags:
        assert np.array_equal(g.tags[t]["faces"], m.tags[t]["faces"]) and np.array_equal(g.tags[t]["ghosts"], m.tags[t]["ghosts"])


@pytest.mark.parametrize("build", BUILDS)
def test_residual_on_native_grid_equals_golden(build):
    """SU2 arrays -> ibis_b200_grid_build -> ibis_b200_create -> compute_dudt, nothing from numpy in between"""
    name = "mixed3d_walls"
    mesh, config, _, prim = cases.build_case(name)
    g = libmod.build_grid(mesh, list(config["grid"]["boundaries"].keys()), lib_path=_lib(build))
    gold = parity.load_golden(name)
    ctx = Context(g, config, lib_path=_lib(build))
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.compute_dudt(fs, dq, True)
    assert np.array_equal(ctx.cq_download(dq)[:g.n_cells], gold["dudt"])
    ctx.close()


@pytest.mark.parametrize("build", BUILDS)
def test_native_grid_larger_box(build):
    """64k jittered hexes: same arrays as the mirror, and the divergence-theorem volumes fill the box"""
    mesh = su2.box_hex(40, jitter=0.2)
    tags = sorted(mesh.markers)
    g = libmod.build_grid(mesh, tags, lib_path=_lib(build))
    m = gridmod.build_grid(mesh, tags)
    for k in ARRAYS:
        assert np.array_equal(np.asarray(getattr(g, k)), np.asarray(getattr(m, k))), k
    assert abs(g.cell_volume.sum() - 1.0) < 1e-12
    assert g.n_faces == 3 * 40 * 40 * 41 and g.n_ghost == 6 * 40 * 40


def test_native_grid_errors():
    L = emu_lib()
    mesh = su2.box_quad(3, 3)
    with pytest.raises(IbisB200Error, match="faces without two cells"):           # a marker left out
        libmod.build_grid(mesh, ["left", "right", "bottom"], lib_path=L)
    bad = su2.box_quad(3, 3)
    t, mv = bad.markers["left"]
    mv = mv.copy()
    mv[0, :2] = [0, 5]                                                             # not an edge of any cell
    bad.markers["left"] = (t, mv)
    with pytest.raises(IbisB200Error, match="does not match any cell face"):
        libmod.build_grid(bad, None, lib_path=L)
    bad = su2.box_quad(3, 3)
    bad.cell_types = bad.cell_types.copy()
    bad.cell_types[2] = 7
    with pytest.raises(IbisB200Error, match="unsupported cell type"):
        libmod.build_grid(bad, None, lib_path=L)
    bad = su2.box_quad(3, 3)
    bad.cell_vertices = bad.cell_vertices.copy()
    bad.cell_vertices[1, 1] = 999
    with pytest.raises(IbisB200Error, match="outside the vertex list"):
        libmod.build_grid(bad, None, lib_path=L)
    hexes = su2.box_hex(2)
    hexes.dim = 2
    with pytest.raises(IbisB200Error, match="unsupported cell type"):              # 3-D cells in a 2-D grid
        libmod.build_grid(hexes, None, lib_path=L)


# ---- ibis_b200_mesh_read_su2 (GridIO::read_su2_grid, grid/grid_io.cpp:218-268) ----
@pytest.mark.parametrize("name", ["mixed3d_walls", "tri2d", "naca_omesh", "box3d_ausmdv"])
def test_native_su2_reader_round_trip(name, tmp_path):
    """the file the reference itself is given (Su2Mesh.write, 17 significant digits) read back natively:
    same vertices bit for bit, same cells, same markers; then file -> grid equals the reference's GridBlock"""
    L = emu_lib()
    mesh, tags = _mesh_of(name)
    path = str(tmp_path / "block_0000.su2")
    mesh.write(path)
    got = libmod.read_su2(path, lib_path=L)
    assert got.dim == mesh.dim
    assert np.array_equal(got.vertices, mesh.vertices)
    assert np.array_equal(got.cell_types, mesh.cell_types) and np.array_equal(got.cell_vertices, mesh.cell_vertices)
    assert list(got.markers) == sorted(mesh.markers)
    for t in mesh.markers:
        assert np.array_equal(got.markers[t][0], mesh.markers[t][0]) and np.array_equal(got.markers[t][1], mesh.markers[t][1])
    ref = su2.Su2Mesh.read(path)
    assert np.array_equal(ref.vertices, got.vertices) and np.array_equal(ref.cell_vertices, got.cell_vertices)
    g = libmod.build_grid(got, tags, lib_path=L)
    gold = parity.load_golden(name)
    for k in ("face_left", "face_right", "face_area", "face_norm", "cell_volume", "cell_centroid", "cell_face_ids"):
        assert np.array_equal(np.asarray(getattr(g, k)), gold["grid_" + k]), k


def test_native_su2_reader_quirks(tmp_path):
    """trailing indices after vertices / elements, unknown lines, a tag split over two MARKER_TAG blocks
    (appended, grid_io.cpp:254-262), tags returned alphabetically; missing file and bad lines are errors"""
    L = emu_lib()
    path = str(tmp_path / "g.su2")
    with open(path, "w") as f:
        f.write("% comment the reader ignores\nNDIME= 2\nNPOIN= 6 \n"
                "0 0 0\n 1.5 0 1\n3 0 2\n0 1e0 3\n1.5 1 4\n  3.0 1.0   5\n"
                "NELEM= 2\n9 0 1 4 3 0\n9 1 2 5 4 1\n"
                "NMARK= 4\nMARKER_TAG= wall\nMARKER_ELEMS= 1\n3 0 1\nMARKER_TAG= in\nMARKER_ELEMS= 1\n3 3 0\n"
                "MARKER_TAG= wall\nMARKER_ELEMS= 3\n3 1 2\n3 4 3\n3 5 4\nMARKER_TAG= out\nMARKER_ELEMS= 1\n3 2 5\n")
    m = libmod.read_su2(path, lib_path=L)
    assert m.dim == 2 and list(m.markers) == ["in", "out", "wall"]
    assert np.array_equal(m.vertices, [[0, 0, 0], [1.5, 0, 0], [3, 0, 0], [0, 1, 0], [1.5, 1, 0], [3, 1, 0]])
    assert np.array_equal(m.cell_vertices[:, :4], [[0, 1, 4, 3], [1, 2, 5, 4]])
    assert np.array_equal(m.markers["wall"][1][:, :2], [[0, 1], [1, 2], [4, 3], [5, 4]])
    g = libmod.build_grid(m, None, lib_path=L)
    assert g.n_faces == 7 and g.n_ghost == 6 and abs(g.cell_volume.sum() - 3.0) < 1e-15
    with pytest.raises(IbisB200Error, match="Could not find"):
        libmod.read_su2(str(tmp_path / "absent.su2"), lib_path=L)
    with open(path, "w") as f:
        f.write("NDIME= 2\nNPOIN= 2\n0 0\n1 0\nNELEM= 1\n11 0 1 2 3 4 5 6 7\n")
    with pytest.raises(IbisB200Error, match="unsupported element"):
        libmod.read_su2(path, lib_path=L)
    with open(path, "w") as f:
        f.write("NDIME= 3\nNPOIN= 5\n0 0 0\n1 0 0\n")
    with pytest.raises(IbisB200Error, match="truncated NPOIN"):
        libmod.read_su2(path, lib_path=L)
    # a line with too few tokens must not borrow numbers from the line below (the reference tokenises line by line)
    with open(path, "w") as f:
        f.write("NDIME= 2\nNPOIN= 3\n0 0\n1\n0 1\nNELEM= 1\n5 0 1 2\n")
    with pytest.raises(IbisB200Error, match="malformed vertex line"):
        libmod.read_su2(path, lib_path=L)
    with open(path, "w") as f:
        f.write("NDIME= 2\nNPOIN= 4\n0 0\n1 0\n1 1\n0 1\nNELEM= 2\n5 0 1\n5 2 3 0\n")
    with pytest.raises(IbisB200Error, match="malformed or unsupported element"):
        libmod.read_su2(path, lib_path=L)
    # a corrupt count fails as a truncated section, not in the allocator
    with open(path, "w") as f:
        f.write("NDIME= 2\nNPOIN= 900000000000000\n0 0\n")
    with pytest.raises(IbisB200Error, match="truncated NPOIN"):
        libmod.read_su2(path, lib_path=L)


def test_empty_and_single_cell_grids():
    """edge sizes (host harness; not yet run on a GPU): an empty mesh builds, creates a context and evaluates
    (nothing); one triangle closed by slip walls has three ghosts, area 1/2, and no mass flux through its walls"""
    from ibis_b200 import config as cfg
    L = emu_lib()
    empty = su2.Su2Mesh(2, np.zeros((0, 3)), np.zeros(0, np.int32), np.zeros((0, 8), np.int64), {})
    g = libmod.build_grid(empty, [], lib_path=L)
    assert (g.n_cells, g.n_faces, g.n_ghost) == (0, 0, 0)
    ctx = Context(g, cfg.make_config(2, {}), lib_path=L)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.compute_dudt(fs, dq, True)
    assert ctx.count_bad_cells(fs) == 0 and ctx.cq_download(dq).shape == (0, 4)
    ctx.close()
    tri = su2.Su2Mesh(2, np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]]), np.array([su2.TRI], np.int32),
                      np.array([[0, 1, 2, -1, -1, -1, -1, -1]]),
                      {"w": (np.array([su2.LINE] * 3, np.int32), np.array([[0, 1, -1, -1], [1, 2, -1, -1], [2, 0, -1, -1]]))})
    g = libmod.build_grid(tri, None, lib_path=L)
    assert (g.n_cells, g.n_faces, g.n_ghost) == (1, 3, 3) and g.cell_volume[0] == 0.5
    assert np.array_equal(np.sort(g.cell_neighbours), [1, 2, 3])
    ctx = Context(g, cfg.make_config(2, {"w": cfg.slip_wall()}, flux_calculator="ausmdv"), lib_path=L)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, cases.uniform_state(g, 2, cases.FS2))
    for allow in (True, False):
        ctx.compute_dudt(fs, dq, allow)
        d = ctx.cq_download(dq)[0]
        assert abs(d[0]) < 1e-9 * 1.225 * 1000.0 / 0.5 and np.isfinite(d).all()
    assert ctx.estimate_dt(fs) > 0
    ctx.close()
