"""The synthetic hex box of BASELINE.json configs[4] built by the C++ closed-form generator
(ibis_b200/host/box_grid.cpp) instead of su2.box_hex + grid.build_grid: same arrays, bit for bit,
in seconds instead of minutes at 64 M cells.  ``BoxGrid`` quacks like ``grid.Grid`` (numpy views
onto the C++ storage, no copies) so that ``lib.Context`` takes it unchanged; ``slab=(k_lo, k_hi)``
gives one rank's z-slab with its two-ring halo and exchange lists."""
import ctypes as C

import numpy as np

from .host import HostLibrary
from .lib import GridDesc, IbisB200Error

TAGS = ["bottom", "east", "north", "south", "top", "west"]


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype) if False else \
        np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(C.addressof(ptr.contents)),
                      dtype=dtype)


class BoxGrid:
    def __init__(self, nx, ny=None, nz=None, lengths=(1.0, 1.0, 1.0), slab=None, host_lib=None, core_lib=None):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        H = HostLibrary.load(host_lib, core_lib).lib
        H.ibis_b200_host_box_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double,
                                                C.c_double, C.c_double, C.c_int, C.c_int]
        H.ibis_b200_host_box_desc.argtypes = [C.c_void_p]
        H.ibis_b200_host_box_desc.restype = C.POINTER(GridDesc)
        H.ibis_b200_host_box_free.argtypes = [C.c_void_p]
        H.ibis_b200_host_box_free.restype = None
        H.ibis_b200_host_box_last_error.restype = C.c_char_p
        i64pp = C.POINTER(C.POINTER(C.c_int64))
        H.ibis_b200_host_box_halo.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int32), i64pp, i64pp,
                                              i64pp, i64pp]
        self._H = H
        self._h = C.c_void_p()
        k_lo, k_hi = slab if slab is not None else (0, nz)
        rc = H.ibis_b200_host_box_create(C.byref(self._h), nx, ny, nz, lengths[0], lengths[1], lengths[2],
                                         int(k_lo), int(k_hi))
        if rc != 0:
            raise IbisB200Error(f"box_create failed ({rc}): {H.ibis_b200_host_box_last_error().decode()}")
        d = H.ibis_b200_host_box_desc(self._h).contents
        self.shape, self.slab = (nx, ny, nz), (int(k_lo), int(k_hi))
        self.dim, self.n_cells, self.n_ghost, self.n_faces = 3, d.n_cells, d.n_ghost, d.n_faces
        self.n_owned = d.n_owned
        nf, nc, nt = d.n_faces, d.n_cells, d.n_cells + d.n_ghost
        self.cell_face_offsets = _view(d.cell_face_offsets, nc + 1, np.int64)
        ncf = int(self.cell_face_offsets[-1])
        self.face_left = _view(d.face_left, nf, np.int64)
        self.face_right = _view(d.face_right, nf, np.int64)
        self.face_area = _view(d.face_area, nf, np.float64)
        self.face_norm = _view(d.face_norm, 3 * nf, np.float64).reshape(nf, 3)
        self.face_tan1 = _view(d.face_tan1, 3 * nf, np.float64).reshape(nf, 3)
        self.face_tan2 = _view(d.face_tan2, 3 * nf, np.float64).reshape(nf, 3)
        self.face_centre = _view(d.face_centre, 3 * nf, np.float64).reshape(nf, 3)
        self.cell_face_ids = _view(d.cell_face_ids, ncf, np.int64)
        self.cell_outsigns = _view(d.cell_outsigns, ncf, np.int32)
        self.cell_neighbours = _view(d.cell_neighbours, ncf, np.int64)
        self.cell_volume = _view(d.cell_volume, nc, np.float64)
        self.cell_centroid = _view(d.cell_centroid, 3 * nt, np.float64).reshape(nt, 3)
        off = _view(d.tag_offsets, 7, np.int64)
        tf = _view(d.tag_faces, int(off[6]), np.int64)
        self.tags = {}
        for t, name in enumerate(TAGS):
            faces = tf[off[t]:off[t + 1]]
            self.tags[name] = dict(faces=faces, ghosts=nc + np.arange(off[t], off[t + 1]))
        # halo lists
        npeers = C.c_int(0)
        pdir = (C.c_int32 * 2)()
        so, sc, ro, rcv = (C.POINTER(C.c_int64)() for _ in range(4))
        H.ibis_b200_host_box_halo(self._h, C.byref(npeers), pdir, C.byref(so), C.byref(sc), C.byref(ro), C.byref(rcv))
        self.n_peers = npeers.value
        self.peer_dir = [pdir[i] for i in range(self.n_peers)]
        self.send_offsets = _view(so, self.n_peers + 1, np.int64)
        self.recv_offsets = _view(ro, self.n_peers + 1, np.int64)
        self.send_cells = _view(sc, int(self.send_offsets[-1]), np.int64)
        self.recv_cells = _view(rcv, int(self.recv_offsets[-1]), np.int64)

    @property
    def n_total(self):
        return self.n_cells + self.n_ghost

    def close(self):
        if self._h:
            self._H.ibis_b200_host_box_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
