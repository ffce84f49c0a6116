"""Host-side grid construction: SU2 description -> the arrays of a reference GridBlock.

Mirror of ``GridBlock<T>::init_grid_block`` (/root/reference/src/grid/grid/grid.h:89-167)
in vectorised numpy.  Numbering and arithmetic follow the reference so that the arrays
are bit-identical to the ones it builds:

* faces are numbered in order of first discovery while looping cells in order and their
  faces in VTK order (grid.h:113-133, grid_io.cpp:298-348); the first discoverer fixes the
  vertex order and hence the orientation (interface.h:187-250);
* ghost cells are appended after the valid cells, boundary tags in alphabetical order
  (nlohmann json object iteration; grid.h:414-433), faces in marker order;
* geometry: face centre = vertex average, area by triangles (geom.h:30-62), tangents /
  normal (interface.h:187-250), cell centroid = vertex average (cell.h:236-254), left/right
  by sign of (x_f - x_c).n (grid.h:186-213), volumes (cell.h:263-303, geom.h:67-83),
  ghost centroids mirrored through the face (grid.h:364-403).

The reference de-duplicates faces through an unordered_map keyed on strings (~17 us per
cell); here it is one lexicographic sort.  Cells may be any mix of tri/quad (2-D) or
tetra/hex/wedge/pyramid (3-D).
"""
import numpy as np

from . import su2 as S

# VTK local face tables (grid_io.cpp:298-348): per cell type, list of local vertex tuples
FACE_TABLE = {
    S.TRI: [(0, 1), (1, 2), (2, 0)],
    S.QUAD: [(0, 1), (1, 2), (2, 3), (3, 0)],
    S.TETRA: [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)],
    S.HEX: [(0, 1, 2, 3), (0, 1, 5, 4), (4, 5, 6, 7), (2, 3, 7, 6), (0, 4, 7, 3), (1, 5, 6, 2)],
    S.WEDGE: [(0, 1, 2), (3, 5, 4), (1, 4, 5, 2), (0, 2, 5, 3), (0, 3, 4, 1)],
    S.PYRAMID: [(0, 3, 2, 1), (2, 3, 4), (0, 4, 3), (0, 1, 4), (1, 2, 4)],
}


class Grid:
    """Plain container of the GridBlock arrays in reference numbering."""

    def __init__(self):
        self.tags = {}

    @property
    def n_total(self):
        return self.n_cells + self.n_ghost

    def as_dict(self):
        d = {k: v for k, v in self.__dict__.items()}
        return d


def _seq_sum(vals, nv, maxn):
    """x = 0.0; for v in vertices: x += pos[v]  (left to right, as the reference loops)."""
    acc = np.zeros(vals.shape[0])
    for k in range(maxn):
        m = nv > k
        acc = np.where(m, acc + np.where(m, vals[:, k], 0.0), acc)
    return acc


def _tri_area(p, a, b, c):
    ab = p[b] - p[a]
    ac = p[c] - p[a]
    cx = ab[:, 1] * ac[:, 2] - ab[:, 2] * ac[:, 1]
    cy = ab[:, 2] * ac[:, 0] - ab[:, 0] * ac[:, 2]
    cz = ab[:, 0] * ac[:, 1] - ab[:, 1] * ac[:, 0]
    return 0.5 * np.sqrt(cx * cx + cy * cy + cz * cz)


def build_grid(mesh, boundary_tags=None):
    """mesh: su2.Su2Mesh.  boundary_tags: the keys of config["grid"]["boundaries"]
    (default: all markers).  Every boundary gets ghost cells (ghost_cells: true is the
    only setting prep.py emits)."""
    dim = mesh.dim
    pos = mesh.vertices
    ctype = mesh.cell_types
    cverts = mesh.cell_vertices
    n_cells = len(ctype)
    if boundary_tags is None:
        boundary_tags = list(mesh.markers.keys())
    boundary_tags = sorted(boundary_tags)

    # ---- (cell, local face) records in reference loop order ----
    nfaces_of = np.zeros(n_cells, np.int64)
    for t, tab in FACE_TABLE.items():
        nfaces_of[ctype == t] = len(tab)
    if np.any(nfaces_of == 0):
        raise ValueError("unsupported cell type in mesh")
    offsets = np.zeros(n_cells + 1, np.int64)
    np.cumsum(nfaces_of, out=offsets[1:])
    n_cf = int(offsets[-1])
    cf_cell = np.repeat(np.arange(n_cells), nfaces_of)
    cf_local = np.arange(n_cf) - offsets[cf_cell]
    cf_verts = -np.ones((n_cf, 4), np.int64)
    cf_nv = np.zeros(n_cf, np.int64)
    for t, tab in FACE_TABLE.items():
        cm = ctype[cf_cell] == t
        if not cm.any():
            continue
        for lf, loc in enumerate(tab):
            m = cm & (cf_local == lf)
            cells_m = cf_cell[m]
            for k, lv in enumerate(loc):
                cf_verts[m, k] = cverts[cells_m, lv]
            cf_nv[m] = len(loc)

    # ---- de-duplicate on the sorted vertex tuple ----
    key = np.sort(cf_verts, axis=1)          # -1 padding sorts first
    order = np.lexsort((np.arange(n_cf), key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    new = np.ones(n_cf, bool)
    new[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    grp = np.cumsum(new) - 1                  # group id in sorted order
    first_rec = order[new]                    # first (lowest record index) of each group
    # face id = rank of the group's first record among all first records
    rank = np.empty(len(first_rec), np.int64)
    rank[np.argsort(first_rec, kind="stable")] = np.arange(len(first_rec))
    cf_face = np.empty(n_cf, np.int64)
    cf_face[order] = rank[grp]
    n_faces = len(first_rec)
    face_first = np.empty(n_faces, np.int64)
    face_first[rank] = first_rec
    fverts = cf_verts[face_first]             # orientation of the first discoverer
    fnv = cf_nv[face_first]
    # sorted key table for marker lookup
    key_sorted = ks[new]
    key_face = rank                           # face id of each sorted unique key

    def lookup(mv):
        k = np.sort(mv, axis=1)
        # search rows of k in key_sorted (lexicographic) via structured view
        a = np.ascontiguousarray(key_sorted).view([("", np.int64)] * 4).ravel()
        b = np.ascontiguousarray(k).view([("", np.int64)] * 4).ravel()
        idx = np.searchsorted(a, b)
        idx = np.clip(idx, 0, len(a) - 1)
        if not np.all(a[idx] == b):
            raise ValueError("boundary marker element does not match any cell face")
        return key_face[idx]

    g = Grid()
    g.dim, g.n_cells, g.n_faces, g.n_vertices = dim, n_cells, n_faces, len(pos)

    # ---- boundaries: ghost ids in alphabetical tag order ----
    n_ghost = 0
    ghost_of_face = {}
    ghost_face_ids, ghost_ids = [], []
    for tag in boundary_tags:
        _, mv = mesh.markers[tag]
        faces = lookup(mv)
        ghosts = n_cells + n_ghost + np.arange(len(faces))
        n_ghost += len(faces)
        g.tags[tag] = dict(faces=faces.astype(np.int64), ghosts=ghosts.astype(np.int64))
        ghost_face_ids.append(faces)
        ghost_ids.append(ghosts)
    g.n_ghost = n_ghost
    n_total = n_cells + n_ghost

    # ---- face geometry ----
    fx = pos[np.maximum(fverts, 0)]           # (nf, 4, 3); padded entries masked below
    nvf = fnv.astype(np.float64)
    centre = np.stack([_seq_sum(fx[:, :, d], fnv, 4) / nvf for d in range(3)], axis=1)
    v0, v1, v2, v3 = (np.maximum(fverts[:, k], 0) for k in range(4))
    area = np.zeros(n_faces)
    is_line, is_tri, is_quad = fnv == 2, fnv == 3, fnv == 4
    if is_line.any():
        d = pos[v1[is_line]] - pos[v0[is_line]]
        area[is_line] = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    if is_tri.any():
        area[is_tri] = _tri_area(pos, v0[is_tri], v1[is_tri], v2[is_tri])
    if is_quad.any():
        area[is_quad] = (_tri_area(pos, v0[is_quad], v1[is_quad], v2[is_quad])
                         + _tri_area(pos, v0[is_quad], v2[is_quad], v3[is_quad]))
    d01 = pos[v1] - pos[v0]
    il = 1.0 / np.sqrt(d01[:, 0] * d01[:, 0] + d01[:, 1] * d01[:, 1] + d01[:, 2] * d01[:, 2])
    tan1 = il[:, None] * d01
    tan2 = np.zeros((n_faces, 3))
    tan2[is_line, 2] = 1.0
    m = ~is_line
    if m.any():
        t2s = pos[v2[m]] - pos[v0[m]]
        t1 = tan1[m]
        proj = t1[:, 0] * t2s[:, 0] + t1[:, 1] * t2s[:, 1] + t1[:, 2] * t2s[:, 2]
        t2 = t2s - proj[:, None] * t1
        il2 = 1.0 / np.sqrt(t2[:, 0] * t2[:, 0] + t2[:, 1] * t2[:, 1] + t2[:, 2] * t2[:, 2])
        tan2[m] = il2[:, None] * t2
    norm = np.stack([tan1[:, 1] * tan2[:, 2] - tan1[:, 2] * tan2[:, 1],
                     tan1[:, 2] * tan2[:, 0] - tan1[:, 0] * tan2[:, 2],
                     tan1[:, 0] * tan2[:, 1] - tan1[:, 1] * tan2[:, 0]], axis=1)

    # ---- cell centroids (valid cells) ----
    cnv = np.array([S.NUM_VERTICES[int(t)] for t in np.unique(ctype)])
    nvc = np.zeros(n_cells, np.int64)
    for t in np.unique(ctype):
        nvc[ctype == t] = S.NUM_VERTICES[int(t)]
    cx = pos[np.maximum(cverts, 0)]           # (nc, 8, 3)
    centroid = np.zeros((n_total, 3))
    for d in range(3):
        centroid[:n_cells, d] = _seq_sum(cx[:, :, d], nvc, 8) / nvc.astype(np.float64)

    # ---- left/right attachment and outsigns ----
    dv = centre[cf_face] - centroid[cf_cell]
    nf_ = norm[cf_face]
    dot = dv[:, 0] * nf_[:, 0] + dv[:, 1] * nf_[:, 1] + dv[:, 2] * nf_[:, 2]
    is_left = dot > 0.0
    outsign = np.where(is_left, 1, -1).astype(np.int32)
    NONE = np.iinfo(np.int64).max
    left = np.full(n_faces, NONE, np.int64)
    right = np.full(n_faces, NONE, np.int64)
    left[cf_face[is_left]] = cf_cell[is_left]
    right[cf_face[~is_left]] = cf_cell[~is_left]
    for faces, ghosts in zip(ghost_face_ids, ghost_ids):
        empty_left = left[faces] == NONE
        left[faces[empty_left]] = ghosts[empty_left]
        right[faces[~empty_left]] = ghosts[~empty_left]
    if np.any(left == NONE) or np.any(right == NONE):
        raise ValueError("grid has faces without two cells: a boundary marker is missing "
                         "or the mesh is non-conforming")

    # ---- volumes ----
    volume = np.zeros(n_cells)
    if dim == 2:
        a, b, c, d_ = (np.maximum(cverts[:, k], 0) for k in range(4))
        tri = ctype == S.TRI
        quad = ctype == S.QUAD
        if tri.any():
            volume[tri] = _tri_area(pos, a[tri], b[tri], c[tri])
        if quad.any():
            volume[quad] = (_tri_area(pos, a[quad], b[quad], c[quad])
                            + _tri_area(pos, a[quad], c[quad], d_[quad]))
    else:
        out = outsign.astype(np.float64)
        cfc, cfn = centre[cf_face], norm[cf_face]
        dotv = (cfc[:, 0] * out * cfn[:, 0] + cfc[:, 1] * out * cfn[:, 1]
                + cfc[:, 2] * out * cfn[:, 2])
        term = dotv * area[cf_face]
        # sequential per-cell accumulation in cell-face order
        maxf = int(nfaces_of.max())
        acc = np.zeros(n_cells)
        for k in range(maxf):
            mk = nfaces_of > k
            idx = offsets[:-1][mk] + k
            acc[mk] = acc[mk] + term[idx]
        volume = acc / 3.0

    # ---- neighbours ----
    neighbours = np.where(left[cf_face] == cf_cell, right[cf_face], left[cf_face])

    # ---- ghost centroids ----
    for tag in boundary_tags:
        faces, ghosts = g.tags[tag]["faces"], g.tags[tag]["ghosts"]
        l, r = left[faces], right[faces]
        valid = np.where(l < n_cells, l, r)
        ghost = np.where(l < n_cells, r, l)
        dxyz = centre[faces] - centroid[valid]
        centroid[ghost] = centre[faces] + dxyz

    g.face_left, g.face_right = left, right
    g.face_area, g.face_norm, g.face_tan1, g.face_tan2, g.face_centre = area, norm, tan1, tan2, centre
    g.face_vertices, g.face_nv = fverts, fnv
    g.cell_face_offsets, g.cell_face_ids = offsets, cf_face
    g.cell_outsigns, g.cell_neighbours = outsign, neighbours
    g.cell_volume, g.cell_centroid = volume, centroid
    g.cell_types = ctype
    return g


def compare_grids(a, b):
    """a: Grid, b: dict from the reference (oracle.refharness.RefCase.grid_arrays).
    Returns the list of array names that are not bit-identical."""
    bad = []
    for k in ("face_left", "face_right", "face_area", "face_norm", "face_tan1", "face_tan2",
              "face_centre", "cell_face_offsets", "cell_face_ids", "cell_outsigns",
              "cell_neighbours", "cell_volume", "cell_centroid"):
        x, y = np.asarray(getattr(a, k)), np.asarray(b[k])
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(k)
    for tag, tv in b["tags"].items():
        if tag not in a.tags or not np.array_equal(a.tags[tag]["faces"], tv["faces"]) \
                or not np.array_equal(a.tags[tag]["ghosts"], tv["ghosts"]):
            bad.append("tag:" + tag)
    return bad
