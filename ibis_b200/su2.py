"""SU2 mesh description: reader, writer and synthetic generators.

The reference reads its grids from SU2 text files (``GridIO::read_su2_grid``,
/root/reference/src/grid/grid/grid_io.cpp:218-268).  This module holds the same
information in numpy form -- vertices, cells as (VTK type, vertex ids) and tagged
boundary markers -- so that a grid can be handed to the reference (through a
file) and to ibis_b200 (in memory) from one description.

The example grids of the reference need gmsh, which is not available; the
generators below produce structured stand-ins with the boundary tags the
example ``job.py`` files use.
"""
import numpy as np

# VTK / SU2 element codes (grid_io.cpp: elem_type_from_vtk_type)
LINE, TRI, QUAD, TETRA, HEX, WEDGE, PYRAMID = 3, 5, 9, 10, 12, 13, 14
NUM_VERTICES = {LINE: 2, TRI: 3, QUAD: 4, TETRA: 4, HEX: 8, WEDGE: 6, PYRAMID: 5}


class Su2Mesh:
    """dim, vertices (nv,3), cell_types (nc,), cell_vertices (nc,8) padded with -1,
    markers: {tag: (types (m,), vertices (m,4) padded with -1)}."""

    def __init__(self, dim, vertices, cell_types, cell_vertices, markers):
        self.dim = int(dim)
        v = np.asarray(vertices, dtype=np.float64)
        if v.shape[1] == 2:
            v = np.concatenate([v, np.zeros((len(v), 1))], axis=1)
        self.vertices = np.ascontiguousarray(v)
        self.cell_types = np.asarray(cell_types, dtype=np.int32)
        cv = np.asarray(cell_vertices, dtype=np.int64)
        if cv.shape[1] < 8:
            cv = np.concatenate([cv, -np.ones((len(cv), 8 - cv.shape[1]), np.int64)], axis=1)
        self.cell_vertices = np.ascontiguousarray(cv)
        self.markers = {}
        for tag, (t, mv) in markers.items():
            mv = np.asarray(mv, dtype=np.int64)
            if mv.shape[1] < 4:
                mv = np.concatenate([mv, -np.ones((len(mv), 4 - mv.shape[1]), np.int64)], axis=1)
            self.markers[tag] = (np.asarray(t, dtype=np.int32), np.ascontiguousarray(mv))

    @property
    def num_cells(self):
        return len(self.cell_types)

    def write(self, path):
        """SU2 text, coordinates with 17 significant digits so std::stod round-trips."""
        with open(path, "w") as f:
            f.write(f"NDIME= {self.dim}\n")
            f.write(f"NPOIN= {len(self.vertices)}\n")
            if self.dim == 2:
                np.savetxt(f, self.vertices[:, :2], fmt="%.17g")
            else:
                np.savetxt(f, self.vertices, fmt="%.17g")
            f.write(f"NELEM= {self.num_cells}\n")
            _write_elems(f, self.cell_types, self.cell_vertices)
            f.write(f"NMARK= {len(self.markers)}\n")
            for tag, (t, mv) in self.markers.items():
                f.write(f"MARKER_TAG= {tag}\n")
                f.write(f"MARKER_ELEMS= {len(t)}\n")
                _write_elems(f, t, mv)

    @staticmethod
    def read(path):
        with open(path) as f:
            lines = [ln.strip() for ln in f]
        i, dim = 0, 2
        verts, ctypes_, cverts, markers = [], [], [], {}

        def val(s):
            return s.split("=")[-1].strip()

        def elem(s, width):
            p = s.split()
            t = int(p[0])
            n = NUM_VERTICES[t]
            return t, [int(x) for x in p[1:1 + n]] + [-1] * (width - n)

        while i < len(lines):
            ln = lines[i]
            if ln.startswith("NDIME"):
                dim = int(val(ln))
            elif ln.startswith("NPOIN"):
                n = int(val(ln).split()[0])
                for k in range(n):
                    p = lines[i + 1 + k].split()
                    verts.append([float(p[0]), float(p[1]), float(p[2]) if dim == 3 else 0.0])
                i += n
            elif ln.startswith("NELEM"):
                n = int(val(ln))
                for k in range(n):
                    t, v = elem(lines[i + 1 + k], 8)
                    ctypes_.append(t)
                    cverts.append(v)
                i += n
            elif ln.startswith("MARKER_TAG"):
                tag = val(ln)
                n = int(val(lines[i + 1]))
                ts, vs = [], []
                for k in range(n):
                    t, v = elem(lines[i + 2 + k], 4)
                    ts.append(t)
                    vs.append(v)
                if tag in markers:
                    markers[tag] = (markers[tag][0] + ts, markers[tag][1] + vs)
                else:
                    markers[tag] = (ts, vs)
                i += n + 1
            i += 1
        markers = {k: (np.array(t, np.int32), np.array(v, np.int64).reshape(-1, 4))
                   for k, (t, v) in markers.items()}
        return Su2Mesh(dim, np.array(verts), np.array(ctypes_, np.int32),
                       np.array(cverts, np.int64).reshape(-1, 8), markers)


def _write_elems(f, types, verts):
    types = np.asarray(types)
    for t in np.unique(types):
        pass
    # fast path: homogeneous element type
    if len(types) and np.all(types == types[0]):
        n = NUM_VERTICES[int(types[0])]
        block = np.concatenate([types[:, None].astype(np.int64), verts[:, :n]], axis=1)
        np.savetxt(f, block, fmt="%d")
        return
    for t, v in zip(types, verts):
        n = NUM_VERTICES[int(t)]
        f.write(" ".join([str(int(t))] + [str(int(x)) for x in v[:n]]) + "\n")


# ----------------------------------------------------------------------------
# generators
# ----------------------------------------------------------------------------
def _jitter(verts, interior_mask, amount, seed):
    if amount <= 0:
        return verts
    rng = np.random.default_rng(seed)
    d = rng.uniform(-amount, amount, size=verts.shape)
    d[~interior_mask] = 0.0
    return verts + d


def box_hex(nx, ny=None, nz=None, lengths=(1.0, 1.0, 1.0), jitter=0.0, seed=0):
    """nx*ny*nz hexes on a box "in unstructured form" (SURVEY.md appendix B.4):
    vertex id = i + (nx+1)(j + (ny+1)k), cells k-major / i-minor, VTK hex ordering,
    markers west/east (x), south/north (y), bottom/top (z).  ``jitter`` (fraction of
    the cell size) perturbs interior vertices to make the cells non-orthogonal."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    lx, ly, lz = lengths
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    vid = (i + (nx + 1) * (j + (ny + 1) * k))
    verts = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    verts[vid.ravel(), 0] = (i.ravel() * lx) / nx
    verts[vid.ravel(), 1] = (j.ravel() * ly) / ny
    verts[vid.ravel(), 2] = (k.ravel() * lz) / nz
    if jitter > 0:
        interior = np.zeros(len(verts), bool)
        m = ((i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz))
        interior[vid[m]] = True
        h = min(lx / nx, ly / ny, lz / nz)
        verts = _jitter(verts, interior[:, None] & np.ones(3, bool)[None, :], jitter * h, seed)

    def v(ii, jj, kk):
        return ii + (nx + 1) * (jj + (ny + 1) * kk)

    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    cells = np.stack([v(ii, jj, kk), v(ii + 1, jj, kk), v(ii + 1, jj + 1, kk), v(ii, jj + 1, kk),
                      v(ii, jj, kk + 1), v(ii + 1, jj, kk + 1), v(ii + 1, jj + 1, kk + 1),
                      v(ii, jj + 1, kk + 1)], axis=1)

    def quads(a, b, fixed, axis):
        # a, b: index ranges of the two in-plane directions
        A, B = np.meshgrid(np.arange(a), np.arange(b), indexing="ij")
        A, B = A.ravel(), B.ravel()
        F = np.full_like(A, fixed)
        if axis == 0:
            q = [v(F, A, B), v(F, A + 1, B), v(F, A + 1, B + 1), v(F, A, B + 1)]
        elif axis == 1:
            q = [v(A, F, B), v(A + 1, F, B), v(A + 1, F, B + 1), v(A, F, B + 1)]
        else:
            q = [v(A, B, F), v(A + 1, B, F), v(A + 1, B + 1, F), v(A, B + 1, F)]
        q = np.stack(q, axis=1)
        return np.full(len(q), QUAD, np.int32), q

    markers = {
        "west": quads(ny, nz, 0, 0), "east": quads(ny, nz, nx, 0),
        "south": quads(nx, nz, 0, 1), "north": quads(nx, nz, ny, 1),
        "bottom": quads(nx, ny, 0, 2), "top": quads(nx, ny, nz, 2),
    }
    return Su2Mesh(3, verts, np.full(len(cells), HEX, np.int32), cells, markers)


def structured_quad(x, y, tags=("left", "right", "bottom", "top"), triangulate=False):
    """Quad (or triangle) cells on a structured (ni+1, nj+1) vertex array x, y.
    Vertex id = i + (ni+1) j, cells j-major / i-minor, counter-clockwise.
    Markers: tags = (i=0, i=ni, j=0, j=nj); a tag of None drops that marker."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ni, nj = x.shape[0] - 1, x.shape[1] - 1
    verts = np.zeros(((ni + 1) * (nj + 1), 3))
    I, J = np.meshgrid(np.arange(ni + 1), np.arange(nj + 1), indexing="ij")
    vid = I + (ni + 1) * J
    verts[vid.ravel(), 0] = x.ravel()
    verts[vid.ravel(), 1] = y.ravel()

    def v(i, j):
        return i + (ni + 1) * j

    jj, ii = np.meshgrid(np.arange(nj), np.arange(ni), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    if triangulate:
        t1 = np.stack([v(ii, jj), v(ii + 1, jj), v(ii + 1, jj + 1)], axis=1)
        t2 = np.stack([v(ii, jj), v(ii + 1, jj + 1), v(ii, jj + 1)], axis=1)
        cells = np.empty((2 * len(t1), 3), np.int64)
        cells[0::2], cells[1::2] = t1, t2
        ctypes_ = np.full(len(cells), TRI, np.int32)
    else:
        cells = np.stack([v(ii, jj), v(ii + 1, jj), v(ii + 1, jj + 1), v(ii, jj + 1)], axis=1)
        ctypes_ = np.full(len(cells), QUAD, np.int32)

    def lines(a, b):
        a, b = np.asarray(a), np.asarray(b)
        return np.full(len(a), LINE, np.int32), np.stack([a, b], axis=1)

    js, is_ = np.arange(nj), np.arange(ni)
    m = {}
    sides = [lines(v(0, js), v(0, js + 1)), lines(v(ni, js), v(ni, js + 1)),
             lines(v(is_, 0), v(is_ + 1, 0)), lines(v(is_, nj), v(is_ + 1, nj))]
    for tag, s in zip(tags, sides):
        if tag is None:
            continue
        if tag in m:
            m[tag] = (np.concatenate([m[tag][0], s[0]]), np.concatenate([m[tag][1], s[1]]))
        else:
            m[tag] = s
    return Su2Mesh(2, verts, ctypes_, cells, m)


def box_quad(nx, ny=None, lengths=(1.0, 1.0), jitter=0.0, seed=0, triangulate=False,
             tags=("left", "right", "bottom", "top")):
    ny = nx if ny is None else ny
    xs = np.arange(nx + 1) * lengths[0] / nx
    ys = np.arange(ny + 1) * lengths[1] / ny
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    if jitter > 0:
        rng = np.random.default_rng(seed)
        h = min(lengths[0] / nx, lengths[1] / ny)
        dx = rng.uniform(-jitter * h, jitter * h, X.shape)
        dy = rng.uniform(-jitter * h, jitter * h, Y.shape)
        dx[0, :] = dx[-1, :] = dx[:, 0] = dx[:, -1] = 0
        dy[0, :] = dy[-1, :] = dy[:, 0] = dy[:, -1] = 0
        X, Y = X + dx, Y + dy
    return structured_quad(X, Y, tags=tags, triangulate=triangulate)


def wedge_2d(ni=60, nj=40, ramp_start=0.2, length=1.0, height=0.6, angle_deg=15.0,
             triangulate=False):
    """Structured stand-in for examples/wedge/2D (supersonic flow over a 15 degree
    ramp).  Tags follow examples/wedge/2D/transient/job.py: inflow, outflow, wall
    (bottom incl. ramp) and top->outflow."""
    xs = np.linspace(0.0, length, ni + 1)
    yb = np.where(xs > ramp_start, (xs - ramp_start) * np.tan(np.radians(angle_deg)), 0.0)
    eta = np.linspace(0.0, 1.0, nj + 1)
    X = np.repeat(xs[:, None], nj + 1, axis=1)
    Y = yb[:, None] + (height - yb[:, None]) * eta[None, :]
    return structured_quad(X, Y, tags=("inflow", "outflow", "wall", "outflow"),
                           triangulate=triangulate)


def flat_plate_2d(ni=60, nj=40, length=1.0, height=0.2, wall_cluster=1.08):
    """Graded stand-in for examples/flat_plate/2D/full_plate: tags inflow, outflow,
    wall (bottom), and top -> outflow (extrapolation)."""
    xs = np.linspace(0.0, length, ni + 1)
    r = wall_cluster
    dy = r ** np.arange(nj)
    ys = np.concatenate([[0.0], np.cumsum(dy)])
    ys = ys / ys[-1] * height
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    return structured_quad(X, Y, tags=("inflow", "outflow", "wall", "outflow"))


def forward_facing_step_2d(h=1.0 / 20):
    """Stand-in for examples/forward_facing_step: 3 x 1 channel with a 0.2 high step
    at x = 0.6, built from quads of size h on a Cartesian lattice with the step
    cells removed.  Tags: inflow, outflow, wall."""
    nx, ny = int(round(3.0 / h)), int(round(1.0 / h))
    sx, sy = int(round(0.6 / h)), int(round(0.2 / h))
    xs = np.arange(nx + 1) * h
    ys = np.arange(ny + 1) * h
    vid = -np.ones((nx + 1, ny + 1), np.int64)
    verts = []
    for j in range(ny + 1):
        for i in range(nx + 1):
            if i > sx and j < sy:
                continue
            vid[i, j] = len(verts)
            verts.append([xs[i], ys[j], 0.0])
    cells = []
    for j in range(ny):
        for i in range(nx):
            if i >= sx and j < sy:
                continue
            cells.append([vid[i, j], vid[i + 1, j], vid[i + 1, j + 1], vid[i, j + 1]])
    inflow = [[vid[0, j], vid[0, j + 1]] for j in range(ny)]
    outflow = [[vid[nx, j], vid[nx, j + 1]] for j in range(sy, ny)]
    wall = [[vid[i, 0], vid[i + 1, 0]] for i in range(sx)]
    wall += [[vid[sx, j], vid[sx, j + 1]] for j in range(sy)]
    wall += [[vid[i, sy], vid[i + 1, sy]] for i in range(sx, nx)]
    wall += [[vid[i, ny], vid[i + 1, ny]] for i in range(nx)]

    def mk(lst):
        return np.full(len(lst), LINE, np.int32), np.array(lst, np.int64)

    return Su2Mesh(2, np.array(verts), np.full(len(cells), QUAD, np.int32),
                   np.array(cells, np.int64), {"inflow": mk(inflow), "outflow": mk(outflow),
                                               "wall": mk(wall)})


def naca0012_omesh(n_around=96, n_radial=48, far=20.0, first=2e-3):
    """O-mesh stand-in for examples/naca0012 (tags: wall, inflow, outflow)."""
    th = np.linspace(0.0, 2 * np.pi, n_around, endpoint=False)
    xc = 0.5 * (1 + np.cos(th))
    t = 0.12
    yt = 5 * t * (0.2969 * np.sqrt(xc) - 0.1260 * xc - 0.3516 * xc ** 2 + 0.2843 * xc ** 3
                  - 0.1036 * xc ** 4)
    yc = np.where(th <= np.pi, yt, -yt)
    xo, yo = 0.5 + far * np.cos(th), far * np.sin(th)
    ratio = (far / first) ** (1.0 / n_radial)
    s = first * (ratio ** np.arange(n_radial + 1) - 1) / (ratio ** n_radial - 1) * (ratio ** n_radial)
    s = s / s[-1]
    X = xc[:, None] + (xo - xc)[:, None] * s[None, :]
    Y = yc[:, None] + (yo - yc)[:, None] * s[None, :]
    na, nr = n_around, n_radial
    verts = np.zeros((na * (nr + 1), 3))
    I, J = np.meshgrid(np.arange(na), np.arange(nr + 1), indexing="ij")
    vid = I + na * J
    verts[vid.ravel(), 0] = X.ravel()
    verts[vid.ravel(), 1] = Y.ravel()

    def v(i, j):
        return (i % na) + na * j

    jj, ii = np.meshgrid(np.arange(nr), np.arange(na), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    # clockwise in (i, j) because theta runs counter-clockwise and j runs outward
    cells = np.stack([v(ii, jj), v(ii, jj + 1), v(ii + 1, jj + 1), v(ii + 1, jj)], axis=1)
    ia = np.arange(na)
    wall = np.stack([v(ia, 0), v(ia + 1, 0)], axis=1)
    outer = np.stack([v(ia, nr), v(ia + 1, nr)], axis=1)
    xm = 0.5 * (verts[outer[:, 0], 0] + verts[outer[:, 1], 0])
    inflow, outflow = outer[xm < 0.5], outer[xm >= 0.5]

    def mk(a):
        return np.full(len(a), LINE, np.int32), a

    return Su2Mesh(2, verts, np.full(len(cells), QUAD, np.int32), cells,
                   {"wall": mk(wall), "inflow": mk(inflow), "outflow": mk(outflow)})


def mixed_3d(n=3, seed=0, jitter=0.15):
    """Small conforming mixed-element 3-D mesh on the unit cube: columns (fixed i, j) of
    hexes, of wedge pairs (split along the 0-2 diagonal of the z-faces) and of hexes cut
    into six pyramids about their centre.  Exercises every entry but the tetrahedron of the
    reference's VTK face tables (grid_io.cpp:298-348)."""
    base = box_hex(n, n, n, jitter=jitter, seed=seed)
    verts = base.vertices
    types, cells, extra = [], [], []
    bottom, top = [], []
    for c, hv in enumerate(base.cell_vertices):
        a = [int(x) for x in hv]
        i, j, k = c % n, (c // n) % n, c // (n * n)
        mode = (i + j) % 3
        if mode == 1:
            types += [WEDGE, WEDGE]
            cells.append([a[0], a[1], a[2], a[4], a[5], a[6], -1, -1])
            cells.append([a[0], a[2], a[3], a[4], a[6], a[7], -1, -1])
            if k == 0:
                bottom += [(TRI, [a[0], a[1], a[2], -1]), (TRI, [a[0], a[2], a[3], -1])]
            if k == n - 1:
                top += [(TRI, [a[4], a[5], a[6], -1]), (TRI, [a[4], a[6], a[7], -1])]
            continue
        if k == 0:
            bottom.append((QUAD, [a[0], a[1], a[2], a[3]]))
        if k == n - 1:
            top.append((QUAD, [a[4], a[5], a[6], a[7]]))
        if mode == 0:
            types.append(HEX)
            cells.append(a)
        else:
            ctr = verts[a].sum(axis=0) / 8.0
            cid = len(verts) + len(extra)
            extra.append(ctr)
            for q in ([a[0], a[3], a[2], a[1]], [a[4], a[5], a[6], a[7]], [a[0], a[1], a[5], a[4]],
                      [a[2], a[3], a[7], a[6]], [a[0], a[4], a[7], a[3]], [a[1], a[2], a[6], a[5]]):
                types.append(PYRAMID)
                cells.append(q + [cid, -1, -1, -1])
    if extra:
        verts = np.concatenate([verts, np.array(extra)], axis=0)
    markers = dict(base.markers)
    for tag, lst in (("bottom", bottom), ("top", top)):
        markers[tag] = (np.array([t for t, _ in lst], np.int32),
                        np.array([v for _, v in lst], np.int64))
    return Su2Mesh(3, verts, np.array(types, np.int32), np.array(cells, np.int64), markers)


def box_tet(n=3, seed=0, jitter=0.1):
    """All-tetrahedra mesh: every hex of box_hex(n) cut into six tets sharing the 0-6
    diagonal (Kuhn split, conforming under translation)."""
    base = box_hex(n, n, n, jitter=jitter, seed=seed)
    paths = [(1, 2), (1, 5), (3, 2), (3, 7), (4, 5), (4, 7)]
    cells = []
    for hv in base.cell_vertices:
        a = [int(x) for x in hv]
        for p, q in paths:
            cells.append([a[0], a[p], a[q], a[6]])
    markers = {}
    # boundary quads -> the two triangles the Kuhn split induces on them (diagonal
    # joins the quad's vertices with the smallest and largest (i+j+k))
    for tag, (t, mv) in base.markers.items():
        tris = []
        for q in mv:
            q = [int(x) for x in q]
            s = base.vertices[q].sum(axis=1)
            lo = int(np.argmin(s))
            r = q[lo:] + q[:lo]
            tris += [[r[0], r[1], r[2], -1], [r[0], r[2], r[3], -1]]
        markers[tag] = (np.full(len(tris), TRI, np.int32), np.array(tris, np.int64))
    return Su2Mesh(3, base.vertices, np.full(len(cells), TETRA, np.int32),
                   np.array(cells, np.int64), markers)
