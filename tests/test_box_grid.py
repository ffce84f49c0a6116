"""The closed-form C++ box generator (ibis_b200/host/box_grid.cpp) against the general numpy
GridBlock mirror (ibis_b200/grid.py, itself bit-identical to the reference's GridBlock on 14
meshes): every array identical, for cubic and non-cubic boxes; slabs: geometry and residual
independent of the decomposition."""
import numpy as np
import pytest

from emu_util import emu_lib, emu_host_lib
from ibis_b200 import grid as gridmod
from ibis_b200 import su2
from ibis_b200.boxgrid import BoxGrid, TAGS

ARRAYS = ("face_left", "face_right", "face_area", "face_norm", "face_tan1", "face_tan2", "face_centre",
          "cell_face_offsets", "cell_face_ids", "cell_outsigns", "cell_neighbours", "cell_volume",
          "cell_centroid")


@pytest.fixture(scope="module")
def libs():
    return emu_lib(), emu_host_lib()


@pytest.mark.parametrize("shape,lengths", [((4, 4, 4), (1.0, 1.0, 1.0)), ((5, 3, 4), (1.0, 1.0, 1.0)),
                                           ((3, 6, 2), (2.0, 1.0, 0.5)), ((1, 1, 1), (1.0, 1.0, 1.0)),
                                           ((7, 2, 9), (1.0, 3.0, 2.0))])
def test_whole_box_identical_to_general_builder(libs, shape, lengths):
    mesh = su2.box_hex(*shape, lengths=lengths)
    g = gridmod.build_grid(mesh, TAGS)
    b = BoxGrid(*shape, lengths=lengths, host_lib=libs[1], core_lib=libs[0])
    assert (b.n_cells, b.n_ghost, b.n_faces, b.n_owned) == (g.n_cells, g.n_ghost, g.n_faces, g.n_cells)
    for k in ARRAYS:
        assert np.array_equal(getattr(b, k), getattr(g, k)), k
    for t in TAGS:
        assert np.array_equal(b.tags[t]["faces"], g.tags[t]["faces"]), t
        assert np.array_equal(b.tags[t]["ghosts"], g.tags[t]["ghosts"]), t
    b.close()


@pytest.mark.parametrize("shape,cuts", [((4, 3, 8), [0, 4, 8]), ((3, 4, 9), [0, 2, 5, 9]), ((5, 5, 12), [0, 3, 6, 9, 12])])
def test_slab_geometry_matches_whole_box(libs, shape, cuts):
    """Every local cell / face of a slab carries the geometry of the same global cell / face."""
    nx, ny, nz = shape
    whole = BoxGrid(*shape, host_lib=libs[1], core_lib=libs[0])
    gcen = {tuple(np.round(c * 1e6).astype(np.int64)): i for i, c in enumerate(whole.cell_centroid[:whole.n_cells])}
    for r in range(len(cuts) - 1):
        s = BoxGrid(*shape, slab=(cuts[r], cuts[r + 1]), host_lib=libs[1], core_lib=libs[0])
        assert s.n_owned == nx * ny * (cuts[r + 1] - cuts[r])
        # local -> global cell map through the (exactly equal) centroids
        l2g = np.array([gcen[tuple(np.round(c * 1e6).astype(np.int64))] for c in s.cell_centroid[:s.n_cells]])
        assert np.array_equal(s.cell_centroid[:s.n_cells], whole.cell_centroid[l2g])
        nfc = int(s.cell_face_offsets[-1]) // 6
        assert np.array_equal(s.cell_volume[:nfc], whole.cell_volume[l2g[:nfc]])
        # per cell-face: same outsign, same face geometry, same neighbour (where the neighbour is held)
        lf = s.cell_face_ids.reshape(nfc, 6)
        gf = whole.cell_face_ids.reshape(whole.n_cells, 6)[l2g[:nfc]]
        assert np.array_equal(s.cell_outsigns.reshape(nfc, 6), whole.cell_outsigns.reshape(-1, 6)[l2g[:nfc]])
        for k in ("face_area", "face_norm", "face_tan1", "face_tan2", "face_centre"):
            assert np.array_equal(getattr(s, k)[lf], getattr(whole, k)[gf]), k
        ln = s.cell_neighbours.reshape(nfc, 6)
        gn = whole.cell_neighbours.reshape(-1, 6)[l2g[:nfc]]
        inner = (ln < s.n_cells)
        assert np.array_equal(l2g[ln[inner]], gn[inner])
        assert np.all(gn[~inner] >= whole.n_cells)                    # ghosts stay ghosts
        assert np.array_equal(s.cell_centroid[ln[~inner]], whole.cell_centroid[gn[~inner]])
        # halo lists: what I receive from a peer is what that peer sends, in the same order
        for p, d in enumerate(s.peer_dir):
            other = BoxGrid(*shape, slab=(cuts[r + d], cuts[r + d + 1]), host_lib=libs[1], core_lib=libs[0])
            q = other.peer_dir.index(-d)
            mine = s.recv_cells[s.recv_offsets[p]:s.recv_offsets[p + 1]]
            theirs = other.send_cells[other.send_offsets[q]:other.send_offsets[q + 1]]
            assert np.array_equal(s.cell_centroid[mine], other.cell_centroid[theirs])
            assert np.all(mine >= s.n_owned) and np.all(theirs < other.n_owned)
            other.close()
        s.close()
    whole.close()
