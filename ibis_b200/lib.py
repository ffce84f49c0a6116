"""ctypes binding of libibis_b200.so (include/ibis_b200.h) and the flattening of the
reference's grid arrays + config.json into its POD structs.

The product path loads ``ibis_b200/lib/libibis_b200.so`` (sm_100a CUDA).  There is no CPU
fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "lib", "libibis_b200.so")

FLUX = {"hanel": 0, "ausmdv": 1, "ldfss2": 2, "rusanov": 3}
LIMITER = {"unlimited": 0, "barth_jespersen": 1}
THERMO = {"rho_p": 0, "rho_T": 1, "rho_u": 2, "p_T": 3}
BC = {"flow_state_copy": 0, "internal_copy": 1, "internal_copy_reflect_normal": 2,
      "internal_vel_copy_reflect": 3, "fix_temperature": 4, "subsonic_inflow": 5,
      "subsonic_outflow": 6, "boundary_layer_profile": 7}
FLUX_ACTION = {"constant_flux": 8}          # build_flux_action (boundary.cpp:482-492)
PHASE = {"pre_reconstruction": 0, "pre_viscous_grad": 1, "post_convective_flux": 2}


class IbisB200Error(RuntimeError):
    pass


class BcAction(C.Structure):
    _fields_ = [("tag", C.c_int32), ("phase", C.c_int32), ("type", C.c_int32),
                ("reserved", C.c_int32), ("params", C.c_double * 8)]


_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)


class GridDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_tags", C.c_int32), ("n_cells", C.c_int64),
                ("n_ghost", C.c_int64), ("n_faces", C.c_int64),
                ("face_left", _i64p), ("face_right", _i64p), ("face_area", _dp),
                ("face_norm", _dp), ("face_tan1", _dp), ("face_tan2", _dp), ("face_centre", _dp),
                ("cell_face_offsets", _i64p), ("cell_face_ids", _i64p), ("cell_outsigns", _i32p),
                ("cell_neighbours", _i64p), ("cell_volume", _dp), ("cell_centroid", _dp),
                ("tag_offsets", _i64p), ("tag_faces", _i64p), ("n_owned", C.c_int64)]


class MeshDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_tags", C.c_int32), ("n_vertices", C.c_int64), ("n_cells", C.c_int64),
                ("vertices", _dp), ("cell_types", _i32p), ("cell_vertices", _i64p),
                ("marker_offsets", _i64p), ("marker_vertices", _i64p)]


class Config(C.Structure):
    _fields_ = [("flux_calculator", C.c_int32), ("reconstruction_order", C.c_int32),
                ("limiter", C.c_int32), ("thermo_interpolator", C.c_int32),
                ("limiter_epsilon", C.c_double), ("ldfss_delta", C.c_double),
                ("viscous", C.c_int32), ("n_bc_actions", C.c_int32),
                ("viscous_signal_factor", C.c_double),
                ("R", C.c_double), ("Cv", C.c_double), ("Cp", C.c_double), ("gamma", C.c_double),
                ("mu0", C.c_double), ("T0", C.c_double), ("Ts", C.c_double), ("Pr", C.c_double),
                ("bc_actions", C.POINTER(BcAction)),
                ("n_bc_tables", C.c_int64), ("bc_tables", C.POINTER(C.c_double))]


def _flow_state_params(fs):
    return [fs["rho"], fs["p"], fs["T"], fs["energy"], fs["vx"], fs["vy"], fs["vz"], 0.0]


def flatten_config(config, tags):
    """config: dict with the reference's config.json schema; tags: boundary tags in the
    order of the grid's tag list (alphabetical).  Unknown names raise like the reference's
    factories do (flux_calc.cpp:20-21, limiter.cpp:16-17, boundary.cpp:474-476)."""
    cf = config["convective_flux"]
    fc = cf["flux_calculator"]
    if fc["type"] not in FLUX:
        raise IbisB200Error("Unknown flux calculator")
    cfgs = Config()
    cfgs.flux_calculator = FLUX[fc["type"]]
    cfgs.ldfss_delta = float(fc.get("delta", 2.0))
    cfgs.reconstruction_order = int(cf["reconstruction_order"])
    cfgs.limiter = LIMITER["barth_jespersen"]
    cfgs.limiter_epsilon = 1e-25
    cfgs.thermo_interpolator = THERMO["rho_u"]
    if cfgs.reconstruction_order > 1:
        lim = cf["limiter"]
        if lim["type"] not in LIMITER:
            raise IbisB200Error("Unknown limiter")
        cfgs.limiter = LIMITER[lim["type"]]
        cfgs.limiter_epsilon = float(lim.get("epsilon", 1e-25))
        if cf["thermo_interpolator"] not in THERMO:
            raise IbisB200Error("Unknown thermo interpolator")
        cfgs.thermo_interpolator = THERMO[cf["thermo_interpolator"]]
    vf = config["viscous_flux"]
    cfgs.viscous = int(bool(vf["enabled"]))
    cfgs.viscous_signal_factor = float(vf["signal_factor"])
    gm = config["gas_model"]
    cfgs.R, cfgs.Cv, cfgs.Cp, cfgs.gamma = (float(gm[k]) for k in ("R", "Cv", "Cp", "gamma"))
    tp = config["transport_properties"]
    if tp["viscosity"]["type"] != "sutherland":
        raise IbisB200Error("Unkown viscosity model")
    if tp["thermal_conductivity"]["type"] != "constant_prandtl_number":
        raise IbisB200Error("Unkown thermal conductivity model")
    cfgs.mu0 = float(tp["viscosity"]["mu_0"])
    cfgs.T0 = float(tp["viscosity"]["T_0"])
    cfgs.Ts = float(tp["viscosity"]["T_s"])
    cfgs.Pr = float(tp["thermal_conductivity"]["Pr"])
    actions, tables = [], []
    bnds = config["grid"]["boundaries"]
    for t, tag in enumerate(tags):
        b = bnds[tag]
        for phase in ("pre_reconstruction", "post_convective_flux", "pre_viscous_grad"):
            flux_phase = phase == "post_convective_flux"
            for act in b.get(phase, []):
                if act["type"] not in (FLUX_ACTION if flux_phase else BC):
                    raise IbisB200Error("Unknown flux action" if flux_phase else "Unknown boundary action")
                a = BcAction()
                a.tag, a.phase, a.type = t, PHASE[phase], (FLUX_ACTION if flux_phase else BC)[act["type"]]
                p = [0.0] * 8
                if act["type"] in ("flow_state_copy", "subsonic_inflow", "constant_flux"):
                    p = _flow_state_params(act["flow_state"])
                elif act["type"] == "fix_temperature":
                    p[0] = float(act["temperature"])
                elif act["type"] == "subsonic_outflow":
                    p[0] = float(act["pressure"])
                elif act["type"] == "boundary_layer_profile":       # boundary.cpp:51-58
                    prof = act["profile"]
                    n = len(prof["height"])
                    if len(prof["v"]) != n or len(prof["T"]) != n:
                        raise IbisB200Error("Cubic spline has different number of x and y points")
                    p[0], p[1], p[2] = float(prof["p"]), float(n), float(len(tables))
                    tables += [float(v) for k in ("height", "v", "T") for v in prof[k]]
                for i in range(8):
                    a.params[i] = p[i]
                actions.append(a)
    arr = (BcAction * max(len(actions), 1))(*actions)
    cfgs.n_bc_actions = len(actions)
    cfgs.bc_actions = C.cast(arr, C.POINTER(BcAction))
    tab = (C.c_double * max(len(tables), 1))(*tables)
    cfgs.n_bc_tables = len(tables)
    cfgs.bc_tables = C.cast(tab, C.POINTER(C.c_double))
    cfgs._keepalive = (arr, tab)
    return cfgs


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def flatten_grid(g, n_owned=0):
    """g: ibis_b200.grid.Grid (or a dict from the reference with the same keys)."""
    get = (lambda k: g[k]) if isinstance(g, dict) else (lambda k: getattr(g, k))
    tags = get("tags")
    names = sorted(tags.keys())
    keep = {}

    def p(name, dtype, ptr):
        a = _as(get(name), dtype)
        keep[name] = a
        return a.ctypes.data_as(ptr)

    d = GridDesc()
    d.dim = int(get("dim"))
    d.n_tags = len(names)
    d.n_cells = int(get("n_cells"))
    d.n_ghost = int(get("n_ghost"))
    d.n_faces = int(get("n_faces"))
    d.face_left = p("face_left", np.int64, _i64p)
    d.face_right = p("face_right", np.int64, _i64p)
    d.face_area = p("face_area", np.float64, _dp)
    d.face_norm = p("face_norm", np.float64, _dp)
    d.face_tan1 = p("face_tan1", np.float64, _dp)
    d.face_tan2 = p("face_tan2", np.float64, _dp)
    d.face_centre = p("face_centre", np.float64, _dp)
    d.cell_face_offsets = p("cell_face_offsets", np.int64, _i64p)
    d.cell_face_ids = p("cell_face_ids", np.int64, _i64p)
    d.cell_outsigns = p("cell_outsigns", np.int32, _i32p)
    d.cell_neighbours = p("cell_neighbours", np.int64, _i64p)
    d.cell_volume = p("cell_volume", np.float64, _dp)
    d.cell_centroid = p("cell_centroid", np.float64, _dp)
    offs = np.zeros(len(names) + 1, np.int64)
    faces = []
    for i, t in enumerate(names):
        faces.append(_as(tags[t]["faces"], np.int64))
        offs[i + 1] = offs[i] + len(faces[-1])
    tf = np.concatenate(faces) if faces else np.zeros(0, np.int64)
    tf = _as(tf, np.int64)
    keep["tag_offsets"], keep["tag_faces"] = offs, tf
    d.tag_offsets = offs.ctypes.data_as(_i64p)
    d.tag_faces = tf.ctypes.data_as(_i64p)
    d.n_owned = int(n_owned)
    d._keepalive = keep
    return d, names


class Library:
    """The loaded shared library with argtypes set."""

    _cache = {}

    @classmethod
    def load(cls, path=None):
        path = os.path.abspath(path or os.environ.get("IBIS_B200_LIB", DEFAULT_LIB))
        if path not in cls._cache:
            cls._cache[path] = cls(path)
        return cls._cache[path]

    def __init__(self, path):
        if not os.path.exists(path):
            raise IbisB200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " (nvcc, sm_100a). ibis_b200 has no CPU fallback.")
        self.path = path
        L = self.lib = C.CDLL(path)
        vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
        ip = C.POINTER(C.c_int)
        sig = {
            "ibis_b200_create": [C.POINTER(vp), C.POINTER(GridDesc), C.POINTER(Config), i],
            "ibis_b200_destroy": [vp], "ibis_b200_sync": [vp],
            "ibis_b200_sizes": [vp, ip, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
            "ibis_b200_fs_create": [vp, i, ip], "ibis_b200_cq_create": [vp, i, ip],
            "ibis_b200_vec_create": [vp, i64, ip], "ibis_b200_free": [vp, i],
            "ibis_b200_fs_upload": [vp, i, vp, vp], "ibis_b200_fs_download": [vp, i, vp, vp],
            "ibis_b200_cq_upload": [vp, i, vp, vp], "ibis_b200_cq_download": [vp, i, vp, vp],
            "ibis_b200_vec_upload": [vp, i, vp], "ibis_b200_vec_download": [vp, i, vp],
            "ibis_b200_copy": [vp, i, i], "ibis_b200_zero": [vp, i],
            "ibis_b200_compute_dudt": [vp, i, i, i],
            "ibis_b200_estimate_dt": [vp, i, C.POINTER(d)],
            "ibis_b200_count_bad_cells": [vp, i, C.POINTER(C.c_uint64)],
            "ibis_b200_cons_to_prim": [vp, i, i], "ibis_b200_prim_to_cons": [vp, i, i],
            "ibis_b200_apply_time_derivative": [vp, i, i, i, d],
            "ibis_b200_l2_norms": [vp, i, vp],
            "ibis_b200_debug_gradients": [vp, i, vp, vp], "ibis_b200_debug_dual_ops": [vp, vp, vp],
            "ibis_b200_debug_division_check": [vp, i64, C.c_uint64, C.POINTER(C.c_uint64)],
            "ibis_b200_set_face_vertices": [vp, vp, i64], "ibis_b200_set_vertex_velocities": [vp, vp],
            "ibis_b200_compute_dudt_moving": [vp, i, i, i, i],
            "ibis_b200_set_cell_vertices": [vp, vp, vp], "ibis_b200_set_vertex_positions": [vp, vp],
            "ibis_b200_move_vertices": [vp, d, i], "ibis_b200_get_vertex_positions": [vp, vp], "ibis_b200_debug_face_velocities": [vp, vp],
            "ibis_b200_rk_step": [vp, i, i, i, vp, vp, d, d, d, C.POINTER(d), C.POINTER(d)],
            "ibis_b200_rk_last_dt": [vp, C.POINTER(d), C.POINTER(d)],
            "ibis_b200_pinned_alloc": [vp, i64, C.POINTER(vp)], "ibis_b200_pinned_free": [vp, vp],
            "ibis_b200_snapshot_upload_begin": [vp, i, vp, vp, vp, vp, vp], "ibis_b200_snapshot_upload_end": [vp, i, i],
            "ibis_b200_snapshot_download_begin": [vp, i, i, vp, vp, vp, vp, vp], "ibis_b200_snapshot_download_end": [vp, i],
            "ibis_b200_rk_residual_norms": [vp, vp], "ibis_b200_rk_distinct_stages": [vp, i],
            "ibis_b200_jfnk_set_dual": [vp, i, i, i],
            "ibis_b200_jfnk_set_vector": [vp, i, i, d, i],
            "ibis_b200_jfnk_eval_rhs": [vp, i, i], "ibis_b200_jfnk_apply_update": [vp, i, i],
            "ibis_b200_vec_dot": [vp, i, i, C.POINTER(d)], "ibis_b200_vec_norm2": [vp, i, C.POINTER(d)],
            "ibis_b200_vec_axpy": [vp, i, i, d], "ibis_b200_vec_scale": [vp, i, i, d],
            "ibis_b200_vec_sub": [vp, i, i, i],
            "ibis_b200_basis_create": [vp, i64, i, ip], "ibis_b200_basis_set_col": [vp, i, i, i],
            "ibis_b200_arnoldi_mgs": [vp, i, i, i, i, vp],
            "ibis_b200_basis_gemv": [vp, i, i, vp, i],
            "ibis_b200_halo_setup": [vp, i, vp, vp, vp, vp, vp],
            "ibis_b200_profile_enable": [vp, i], "ibis_b200_profile_read": [vp, vp, vp],
            "ibis_b200_comm_init": [vp, vp, i, i], "ibis_b200_nccl_unique_id": [vp],
            "ibis_b200_halo_p2p_export": [vp, vp, vp], "ibis_b200_halo_p2p_import": [vp, i, vp, i64, i64, i],
            "ibis_b200_halo_p2p_enable": [vp, i],
            "ibis_b200_flow_read_native": [vp, i, vp, vp, vp, vp, vp],
            "ibis_b200_flow_write_native": [vp, i, vp, vp, vp, vp, vp],
            "ibis_b200_grid_build": [C.POINTER(vp), C.POINTER(MeshDesc), i], "ibis_b200_grid_destroy": [vp],
            "ibis_b200_mesh_read_su2": [C.POINTER(vp), C.c_char_p], "ibis_b200_mesh_destroy": [vp],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        L.ibis_b200_last_error.argtypes = [vp]
        L.ibis_b200_last_error.restype = C.c_char_p
        L.ibis_b200_launch_count.argtypes = [vp]
        L.ibis_b200_launch_count.restype = C.c_int64
        L.ibis_b200_stream.argtypes = [vp]
        L.ibis_b200_stream.restype = vp
        L.ibis_b200_grid_desc_of.argtypes = [vp]
        L.ibis_b200_grid_desc_of.restype = C.POINTER(GridDesc)
        L.ibis_b200_mesh_desc_of.argtypes = [vp]
        L.ibis_b200_mesh_desc_of.restype = C.POINTER(MeshDesc)
        L.ibis_b200_mesh_tag.argtypes = [vp, i]
        L.ibis_b200_mesh_tag.restype = C.c_char_p
        for name in ("ibis_b200_grid_face_vertices", "ibis_b200_grid_tag_ghosts"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = _i64p

    EXPORTS = ["ibis_b200_last_error", "ibis_b200_launch_count", "ibis_b200_stream", "ibis_b200_grid_desc_of",
               "ibis_b200_grid_face_vertices", "ibis_b200_grid_tag_ghosts", "ibis_b200_mesh_desc_of", "ibis_b200_mesh_tag"]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def read_su2(path, lib_path=None):
    """ibis_b200_mesh_read_su2: SU2 text file -> su2.Su2Mesh (markers in alphabetical tag order)."""
    from . import su2
    L = Library.load(lib_path)
    h = C.c_void_p()
    rc = L.lib.ibis_b200_mesh_read_su2(C.byref(h), os.fsencode(path))
    if rc != 0:
        raise IbisB200Error(f"ibis_b200_mesh_read_su2 failed ({rc}): {L.lib.ibis_b200_last_error(None).decode()}")
    try:
        d = L.lib.ibis_b200_mesh_desc_of(h).contents
        nv, nc, nt = int(d.n_vertices), int(d.n_cells), int(d.n_tags)

        def arr(ptr, shape, dtype):
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy() if n else np.zeros(shape, dtype)

        offs = arr(d.marker_offsets, (nt + 1,), np.int64)
        mv = arr(d.marker_vertices, (int(offs[-1]), 4), np.int64)
        line_tri_quad = np.array([0, 0, su2.LINE, su2.TRI, su2.QUAD], np.int32)
        markers = {}
        for k in range(nt):
            v = mv[offs[k]:offs[k + 1]]
            markers[L.lib.ibis_b200_mesh_tag(h, k).decode()] = (line_tri_quad[(v >= 0).sum(axis=1)], v)
        return su2.Su2Mesh(int(d.dim), arr(d.vertices, (nv, 3), np.float64), arr(d.cell_types, (nc,), np.int32),
                           arr(d.cell_vertices, (nc, 8), np.int64), markers)
    finally:
        L.lib.ibis_b200_mesh_destroy(h)


def build_grid(mesh, boundary_tags=None, device=-1, lib_path=None):
    """ibis_b200_grid_build: su2.Su2Mesh -> grid.Grid with the arrays of the reference GridBlock in the
    reference's numbering (same result as the numpy mirror grid.build_grid, which the tests hold it to)."""
    from .grid import Grid
    L = Library.load(lib_path)
    tags = sorted(mesh.markers.keys() if boundary_tags is None else boundary_tags)
    mv = [_as(mesh.markers[t][1], np.int64) for t in tags]
    offs = np.zeros(len(tags) + 1, np.int64)
    offs[1:] = np.cumsum([len(a) for a in mv])
    mverts = _as(np.concatenate(mv) if mv else np.zeros((0, 4)), np.int64)
    pos, ctype, cverts = _as(mesh.vertices, np.float64), _as(mesh.cell_types, np.int32), _as(mesh.cell_vertices, np.int64)
    md = MeshDesc(int(mesh.dim), len(tags), len(pos), len(ctype), pos.ctypes.data_as(_dp), ctype.ctypes.data_as(_i32p),
                  cverts.ctypes.data_as(_i64p), offs.ctypes.data_as(_i64p), mverts.ctypes.data_as(_i64p))
    h = C.c_void_p()
    rc = L.lib.ibis_b200_grid_build(C.byref(h), C.byref(md), int(device))
    if rc != 0:
        raise IbisB200Error(f"ibis_b200_grid_build failed ({rc}): {L.lib.ibis_b200_last_error(None).decode()}")
    try:
        d = L.lib.ibis_b200_grid_desc_of(h).contents
        nc, ng, nf = int(d.n_cells), int(d.n_ghost), int(d.n_faces)

        def arr(ptr, shape):
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy() if n else np.zeros(shape, ptr._type_)

        g = Grid()
        g.dim, g.n_cells, g.n_ghost, g.n_faces, g.n_vertices = int(d.dim), nc, ng, nf, len(pos)
        g.cell_face_offsets = arr(d.cell_face_offsets, (nc + 1,))
        nr = int(g.cell_face_offsets[-1])
        g.face_left, g.face_right = arr(d.face_left, (nf,)), arr(d.face_right, (nf,))
        g.face_area = arr(d.face_area, (nf,))
        for k in ("norm", "tan1", "tan2", "centre"):
            setattr(g, "face_" + k, arr(getattr(d, "face_" + k), (nf, 3)))
        g.cell_face_ids, g.cell_outsigns = arr(d.cell_face_ids, (nr,)), arr(d.cell_outsigns, (nr,))
        g.cell_neighbours = arr(d.cell_neighbours, (nr,))
        g.cell_volume, g.cell_centroid = arr(d.cell_volume, (nc,)), arr(d.cell_centroid, (nc + ng, 3))
        g.face_vertices = arr(L.lib.ibis_b200_grid_face_vertices(h), (nf, 4))
        g.face_nv = (g.face_vertices >= 0).sum(axis=1)
        tf, tg = arr(d.tag_faces, (ng,)), arr(L.lib.ibis_b200_grid_tag_ghosts(h), (ng,))
        for i, t in enumerate(tags):
            g.tags[t] = dict(faces=tf[offs[i]:offs[i + 1]], ghosts=tg[offs[i]:offs[i + 1]])
        g.cell_types = ctype
        return g
    finally:
        L.lib.ibis_b200_grid_destroy(h)


class Context:
    """One ibis_b200_ctx: device grid + configuration + buffers.  Thin, 1:1 with the C ABI."""

    def __init__(self, grid, config, device=-1, lib_path=None, n_owned=0):
        self.L = Library.load(lib_path)
        self.lib = self.L.lib
        gd, self.tags = flatten_grid(grid, n_owned or getattr(grid, "n_owned", 0))
        cf = flatten_config(config, self.tags)
        self.dim = gd.dim
        self.n_cons = gd.dim + 2
        self.n_cells, self.n_ghost, self.n_faces = gd.n_cells, gd.n_ghost, gd.n_faces
        self.n_total = gd.n_cells + gd.n_ghost
        self.h = C.c_void_p()
        rc = self.lib.ibis_b200_create(C.byref(self.h), C.byref(gd), C.byref(cf), int(device))
        if rc != 0:
            raise IbisB200Error(f"ibis_b200_create failed ({rc}): "
                                f"{self.lib.ibis_b200_last_error(None).decode()}")

    def close(self):
        if self.h:
            for p in getattr(self, "_pinned", []):
                self.lib.ibis_b200_pinned_free(self.h, p)
            self._pinned = []
            self.lib.ibis_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sizes(self):
        """GridBlock::dim / num_cells / num_ghost_cells / num_interfaces as the context sees them."""
        dim, nc, ng, nf = C.c_int(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.ibis_b200_sizes(self.h, C.byref(dim), C.byref(nc), C.byref(ng), C.byref(nf)), "sizes")
        return dim.value, nc.value, ng.value, nf.value

    def _ck(self, rc, what):
        if rc != 0:
            raise IbisB200Error(f"{what} failed ({rc}): {self.lib.ibis_b200_last_error(self.h).decode()}")

    # buffers
    def fs_create(self, dual=False):
        h = C.c_int(0)
        self._ck(self.lib.ibis_b200_fs_create(self.h, int(dual), C.byref(h)), "fs_create")
        return h.value

    def cq_create(self, dual=False):
        h = C.c_int(0)
        self._ck(self.lib.ibis_b200_cq_create(self.h, int(dual), C.byref(h)), "cq_create")
        return h.value

    def vec_create(self, n):
        h = C.c_int(0)
        self._ck(self.lib.ibis_b200_vec_create(self.h, int(n), C.byref(h)), "vec_create")
        return h.value

    def basis_create(self, n, n_cols):
        h = C.c_int(0)
        self._ck(self.lib.ibis_b200_basis_create(self.h, int(n), int(n_cols), C.byref(h)), "basis_create")
        return h.value

    def free(self, h):
        self._ck(self.lib.ibis_b200_free(self.h, h), "free")

    def fs_upload(self, h, prim, prim_du=None):
        a = _as(prim, np.float64).reshape(self.n_total, 7)
        b = None if prim_du is None else _as(prim_du, np.float64).reshape(self.n_total, 7)
        self._ck(self.lib.ibis_b200_fs_upload(self.h, h, _ptr(a), _ptr(b)), "fs_upload")

    def fs_download(self, h, dual=False):
        a = np.zeros((self.n_total, 7))
        b = np.zeros((self.n_total, 7)) if dual else None
        self._ck(self.lib.ibis_b200_fs_download(self.h, h, _ptr(a), _ptr(b)), "fs_download")
        return (a, b) if dual else a

    def cq_upload(self, h, cons, cons_du=None):
        a = _as(cons, np.float64).reshape(self.n_total, self.n_cons)
        b = None if cons_du is None else _as(cons_du, np.float64).reshape(self.n_total, self.n_cons)
        self._ck(self.lib.ibis_b200_cq_upload(self.h, h, _ptr(a), _ptr(b)), "cq_upload")

    def cq_download(self, h, dual=False):
        a = np.zeros((self.n_total, self.n_cons))
        b = np.zeros((self.n_total, self.n_cons)) if dual else None
        self._ck(self.lib.ibis_b200_cq_download(self.h, h, _ptr(a), _ptr(b)), "cq_download")
        return (a, b) if dual else a

    def vec_upload(self, h, x):
        a = _as(x, np.float64)
        self._ck(self.lib.ibis_b200_vec_upload(self.h, h, _ptr(a)), "vec_upload")

    def vec_download(self, h, n):
        a = np.zeros(int(n))
        self._ck(self.lib.ibis_b200_vec_download(self.h, h, _ptr(a)), "vec_download")
        return a

    def flow_read_native(self, h, T, p, vx, vy, vz=None):
        """The reference's native_binary snapshot arrays (one f64 array per variable over the valid
        cells) -> FlowStates, with update_thermo_from_pT; arrays may be numpy arrays or raw pointers."""
        keep = [x if isinstance(x, int) or x is None else _as(x, np.float64) for x in (T, p, vx, vy, vz)]
        a = [C.c_void_p(x) if isinstance(x, int) else _ptr(x) for x in keep]
        self._ck(self.lib.ibis_b200_flow_read_native(self.h, h, *a), "flow_read_native")

    def flow_write_native(self, h, out=None):
        """FlowStates -> (T, p, vx, vy[, vz]); `out` = five raw pointers to write through instead."""
        if out is not None:
            self._ck(self.lib.ibis_b200_flow_write_native(self.h, h, *[C.c_void_p(x) if x else None for x in out]),
                     "flow_write_native")
            return None
        dim, n, _, _ = self.sizes()
        arrs = [np.empty(n) for _ in range(5 if dim == 3 else 4)]
        ptrs = [_ptr(x) for x in arrs] + ([None] if dim == 2 else [])
        self._ck(self.lib.ibis_b200_flow_write_native(self.h, h, *ptrs), "flow_write_native")
        return arrs

    # snapshot ring (include/ibis_b200.h): native_binary arrays in and out while the device computes
    def pinned_array(self, shape):
        """a page-locked float64 host array owned by this context (freed by close())"""
        n = int(np.prod(shape))
        p = C.c_void_p()
        self._ck(self.lib.ibis_b200_pinned_alloc(self.h, n * 8, C.byref(p)), "pinned_alloc")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_double * n).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def _rows(self, a):
        """(5 or 4, n_cells) array -> the five row pointers T p vx vy vz (vz NULL in 2-D)"""
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape[1] == self.n_cells
        base, row = a.ctypes.data, self.n_cells * 8
        ptrs = [C.c_void_p(base + q * row) for q in range(a.shape[0])]
        return ptrs + [None] * (5 - len(ptrs))

    def snapshot_upload_begin(self, slot, arrays):
        self._ck(self.lib.ibis_b200_snapshot_upload_begin(self.h, int(slot), *self._rows(arrays)), "snapshot_upload_begin")

    def snapshot_upload_end(self, slot, fs):
        self._ck(self.lib.ibis_b200_snapshot_upload_end(self.h, int(slot), fs), "snapshot_upload_end")

    def snapshot_download_begin(self, slot, fs, arrays):
        self._ck(self.lib.ibis_b200_snapshot_download_begin(self.h, int(slot), fs, *self._rows(arrays)), "snapshot_download_begin")

    def snapshot_download_end(self, slot):
        self._ck(self.lib.ibis_b200_snapshot_download_end(self.h, int(slot)), "snapshot_download_end")

    def copy(self, dst, src):
        self._ck(self.lib.ibis_b200_copy(self.h, dst, src), "copy")

    def zero(self, h):
        self._ck(self.lib.ibis_b200_zero(self.h, h), "zero")

    def sync(self):
        self._ck(self.lib.ibis_b200_sync(self.h), "sync")

    def profile_enable(self, on=True):
        self._ck(self.lib.ibis_b200_profile_enable(self.h, int(on)), "profile_enable")

    def profile_read(self):
        ms = np.zeros(4)
        n = np.zeros(4, np.int64)
        self._ck(self.lib.ibis_b200_profile_read(self.h, _ptr(ms), _ptr(n)), "profile_read")
        return ms, n

    def stream_ptr(self):
        return int(self.lib.ibis_b200_stream(self.h) or 0)

    def fs_upload_ptr(self, h, ptr):
        self._ck(self.lib.ibis_b200_fs_upload(self.h, h, C.c_void_p(ptr), None), "fs_upload")

    def fs_download_ptr(self, h, ptr):
        self._ck(self.lib.ibis_b200_fs_download(self.h, h, C.c_void_p(ptr), None), "fs_download")

    def launch_count(self):
        return int(self.lib.ibis_b200_launch_count(self.h))

    # hot path
    def compute_dudt(self, fs, dudt, allow_reconstruction=True):
        self._ck(self.lib.ibis_b200_compute_dudt(self.h, fs, dudt, int(allow_reconstruction)), "compute_dudt")

    def estimate_dt(self, fs):
        v = C.c_double(0)
        self._ck(self.lib.ibis_b200_estimate_dt(self.h, fs, C.byref(v)), "estimate_dt")
        return v.value

    def count_bad_cells(self, fs):
        v = C.c_uint64(0)
        self._ck(self.lib.ibis_b200_count_bad_cells(self.h, fs, C.byref(v)), "count_bad_cells")
        return int(v.value)

    def cons_to_prim(self, cq, fs):
        self._ck(self.lib.ibis_b200_cons_to_prim(self.h, cq, fs), "cons_to_prim")

    def prim_to_cons(self, cq, fs):
        self._ck(self.lib.ibis_b200_prim_to_cons(self.h, cq, fs), "prim_to_cons")

    def apply_time_derivative(self, u0, u1, dudt, dt):
        self._ck(self.lib.ibis_b200_apply_time_derivative(self.h, u0, u1, dudt, float(dt)), "apply_time_derivative")

    def l2_norms(self, cq):
        out = np.zeros(6)
        self._ck(self.lib.ibis_b200_l2_norms(self.h, cq, _ptr(out)), "l2_norms")
        return out

    def debug_gradients(self, var):
        g = np.zeros((self.n_cells, 3))
        l = np.zeros(self.n_cells)
        self._ck(self.lib.ibis_b200_debug_gradients(self.h, var, _ptr(g), _ptr(l)), "debug_gradients")
        return g, l

    # moving grid
    def set_face_vertices(self, face_vertices, n_vertices):
        a = _as(face_vertices, np.int64)
        self._ck(self.lib.ibis_b200_set_face_vertices(self.h, _ptr(a), int(n_vertices)), "set_face_vertices")

    def set_vertex_velocities(self, vertex_vel):
        a = _as(vertex_vel, np.float64)
        self._ck(self.lib.ibis_b200_set_vertex_velocities(self.h, _ptr(a)), "set_vertex_velocities")

    def set_cell_vertices(self, cell_types, cell_vertices):
        t, v = _as(cell_types, np.int32), _as(cell_vertices, np.int64)
        self._ck(self.lib.ibis_b200_set_cell_vertices(self.h, _ptr(t), _ptr(v)), "set_cell_vertices")

    def set_vertex_positions(self, positions):
        a = _as(positions, np.float64)
        self._n_vertices = len(a)
        self._ck(self.lib.ibis_b200_set_vertex_positions(self.h, _ptr(a)), "set_vertex_positions")

    def move_vertices(self, scale, recompute_geometry=True):
        self._ck(self.lib.ibis_b200_move_vertices(self.h, float(scale), int(recompute_geometry)), "move_vertices")

    def get_vertex_positions(self):
        out = np.zeros((self._n_vertices, 3))
        self._ck(self.lib.ibis_b200_get_vertex_positions(self.h, _ptr(out)), "get_vertex_positions")
        return out

    def compute_dudt_moving(self, fs, cq, dudt, allow_reconstruction=True):
        self._ck(self.lib.ibis_b200_compute_dudt_moving(self.h, fs, cq, dudt, int(allow_reconstruction)),
                 "compute_dudt_moving")

    def debug_face_velocities(self):
        out = np.zeros((self.n_faces, 3))
        self._ck(self.lib.ibis_b200_debug_face_velocities(self.h, _ptr(out)), "debug_face_velocities")
        return out

    def debug_division_check(self, n, seed=1):
        """-> number of quotients of the paired / shared device divisions that differ from the plain IEEE division"""
        bad = C.c_uint64(0)
        self._ck(self.lib.ibis_b200_debug_division_check(self.h, int(n), int(seed), C.byref(bad)), "debug_division_check")
        return int(bad.value)

    def debug_dual_ops(self, x, y, r):
        """-> (22, 2) array of (re, du): see ibis_b200_debug_dual_ops in include/ibis_b200.h"""
        a, out = _as([x[0], x[1], y[0], y[1], r], np.float64), np.zeros(44)
        self._ck(self.lib.ibis_b200_debug_dual_ops(self.h, _ptr(a), _ptr(out)), "debug_dual_ops")
        return out.reshape(22, 2)

    def rk_step(self, fs, cq, a_rows, b, cfl, dt_prev, t_remaining):
        a = np.array([x for row in a_rows for x in row], dtype=np.float64)
        if a.size == 0:
            a = np.zeros(1)
        bb = _as(b, np.float64)
        dt, dts = C.c_double(0), C.c_double(0)
        self._ck(self.lib.ibis_b200_rk_step(self.h, fs, cq, len(bb), _ptr(a), _ptr(bb), float(cfl),
                                            float(dt_prev), float(t_remaining), C.byref(dt), C.byref(dts)),
                 "rk_step")
        return dt.value, dts.value

    def rk_step_async(self, fs, cq, a_rows, b, cfl, t_remaining):
        """one RungeKutta::take_step enqueued on the stream, nothing read back: dt_prev is the previous step's dt on
        the device; rk_last_dt() fetches (dt, dt_stable) of the last step"""
        a = np.array([x for row in a_rows for x in row], dtype=np.float64)
        if a.size == 0:
            a = np.zeros(1)
        bb = _as(b, np.float64)
        self._ck(self.lib.ibis_b200_rk_step(self.h, fs, cq, len(bb), _ptr(a), _ptr(bb), float(cfl),
                                            float("nan"), float(t_remaining), None, None), "rk_step")

    def rk_last_dt(self):
        dt, dts = C.c_double(0), C.c_double(0)
        self._ck(self.lib.ibis_b200_rk_last_dt(self.h, C.byref(dt), C.byref(dts)), "rk_last_dt")
        return dt.value, dts.value

    def rk_residual_norms(self):
        out = np.zeros(6)
        self._ck(self.lib.ibis_b200_rk_residual_norms(self.h, _ptr(out)), "rk_residual_norms")
        return out

    # domain decomposition
    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        self._ck(self.lib.ibis_b200_nccl_unique_id(buf), "nccl_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.ibis_b200_comm_init(self.h, buf, int(rank), int(n_ranks)), "comm_init")

    def halo_setup(self, peer_ranks, send_offsets, send_cells, recv_offsets, recv_cells):
        pr = _as(peer_ranks, np.int32)
        so, sc = _as(send_offsets, np.int64), _as(send_cells, np.int64)
        ro, rc_ = _as(recv_offsets, np.int64), _as(recv_cells, np.int64)
        self._ck(self.lib.ibis_b200_halo_setup(self.h, len(pr), _ptr(pr), _ptr(so), _ptr(sc), _ptr(ro), _ptr(rc_)),
                 "halo_setup")

    def halo_p2p_export(self):
        """(64-byte CUDA IPC handle of this rank's receive allocation, its receive count)."""
        buf = (C.c_char * 64)()
        n = C.c_int64(0)
        self._ck(self.lib.ibis_b200_halo_p2p_export(self.h, buf, C.byref(n)), "halo_p2p_export")
        return bytes(buf), int(n.value)

    def halo_p2p_import(self, peer_index, handle, peer_n_recv, my_offset_in_peer, my_slot_in_peer):
        buf = (C.c_char * 64).from_buffer_copy(handle)
        self._ck(self.lib.ibis_b200_halo_p2p_import(self.h, int(peer_index), buf, int(peer_n_recv),
                                                    int(my_offset_in_peer), int(my_slot_in_peer)), "halo_p2p_import")

    def halo_p2p_enable(self, on=True):
        self._ck(self.lib.ibis_b200_halo_p2p_enable(self.h, int(on)), "halo_p2p_enable")

    def connect(self, dist, rank, world, peer_ranks, send_offsets, send_cells, recv_offsets, recv_cells, p2p=True):
        """The host plumbing of one rank of a decomposed run over torch.distributed: NCCL
        communicator for the scalar reductions, exchange lists, and (p2p) the CUDA IPC handles of
        the peer-memory halo exchange handed round with all_gather_object."""
        ids = [self.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.comm_init(ids[0], rank, world)
        peer_ranks = [int(r) for r in peer_ranks]
        self.halo_setup(peer_ranks, send_offsets, send_cells, recv_offsets, recv_cells)
        if not p2p:
            return "nccl"
        handle, n_recv = self.halo_p2p_export()
        mine = (handle, n_recv, peer_ranks, [int(o) for o in recv_offsets])
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        for p, q in enumerate(peer_ranks):
            q_handle, q_n_recv, q_peers, q_recv_off = everyone[q]
            slot = q_peers.index(rank)
            self.halo_p2p_import(p, q_handle, q_n_recv, q_recv_off[slot], slot)
        dist.barrier()
        self.halo_p2p_enable(True)
        return "p2p"

    # JFNK
    def jfnk_set_dual(self, cq, vec, cq_tmp):
        self._ck(self.lib.ibis_b200_jfnk_set_dual(self.h, cq, vec, cq_tmp), "jfnk_set_dual")

    def jfnk_set_vector(self, residual, vec, dt_star, result):
        self._ck(self.lib.ibis_b200_jfnk_set_vector(self.h, residual, vec, float(dt_star), result), "jfnk_set_vector")

    def jfnk_eval_rhs(self, residual, rhs):
        self._ck(self.lib.ibis_b200_jfnk_eval_rhs(self.h, residual, rhs), "jfnk_eval_rhs")

    def jfnk_apply_update(self, cq, du):
        self._ck(self.lib.ibis_b200_jfnk_apply_update(self.h, cq, du), "jfnk_apply_update")

    # vectors
    def vec_dot(self, a, b):
        v = C.c_double(0)
        self._ck(self.lib.ibis_b200_vec_dot(self.h, a, b, C.byref(v)), "vec_dot")
        return v.value

    def vec_norm2(self, a):
        v = C.c_double(0)
        self._ck(self.lib.ibis_b200_vec_norm2(self.h, a, C.byref(v)), "vec_norm2")
        return v.value

    def vec_axpy(self, y, x, s):
        self._ck(self.lib.ibis_b200_vec_axpy(self.h, y, x, float(s)), "vec_axpy")

    def vec_scale(self, x, result, s):
        self._ck(self.lib.ibis_b200_vec_scale(self.h, x, result, float(s)), "vec_scale")

    def vec_sub(self, a, b, result):
        self._ck(self.lib.ibis_b200_vec_sub(self.h, a, b, result), "vec_sub")

    def basis_set_col(self, basis, col, vec):
        self._ck(self.lib.ibis_b200_basis_set_col(self.h, basis, col, vec), "basis_set_col")

    def arnoldi_mgs(self, basis, j, w, v):
        h = np.zeros(j + 2)
        self._ck(self.lib.ibis_b200_arnoldi_mgs(self.h, basis, j, w, v, _ptr(h)), "arnoldi_mgs")
        return h

    def basis_gemv(self, basis, m, y, out):
        yy = _as(y, np.float64)
        self._ck(self.lib.ibis_b200_basis_gemv(self.h, basis, int(m), _ptr(yy), out), "basis_gemv")
