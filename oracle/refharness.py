"""oracle/refharness.py -- TEST INFRASTRUCTURE, not product code.

ctypes/numpy wrapper around oracle/_ref/libibis_ref.so, i.e. around the
UNMODIFIED reference (wattrg/ibis) compiled by oracle/ref_build/Makefile, plus
a runner for the reference's own ``prep.py`` (job.py -> config/config.json).

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and the
reference / cpu_baseline legs of bench.py may import this module.  The product
path (ibis_b200/) never does.
"""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libibis_ref.so")
REFERENCE_ROOT = os.environ.get("IBIS_REFERENCE_ROOT", "/root/reference")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(REF_LIB)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefLib:
    """Process-wide handle on the reference library (Kokkos initialised once)."""

    _instance = None

    @classmethod
    def get(cls, num_threads=0):
        if cls._instance is None:
            cls._instance = cls(num_threads)
        return cls._instance

    def __init__(self, num_threads=0):
        if not available():
            raise RuntimeError(
                f"{REF_LIB} missing: run `make -C oracle/ref_build -j8` (needs /root/reference)")
        self.lib = C.CDLL(REF_LIB, mode=C.RTLD_GLOBAL)
        L = self.lib
        L.ref_init.argtypes = [C.c_int]
        L.ref_num_threads.restype = C.c_int
        L.ref_case_create.restype = C.c_void_p
        L.ref_case_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_case_destroy.argtypes = [C.c_void_p]
        L.ref_case_sizes.argtypes = [C.c_void_p, _i64p]
        L.ref_grid_faces.argtypes = [C.c_void_p, _i64p, _i64p, _dp, _dp, _dp, _dp, _dp]
        L.ref_grid_cells.argtypes = [C.c_void_p, _i64p, _i64p, _i32p, _i64p, _dp, _dp]
        L.ref_grid_tag.restype = C.c_int64
        L.ref_grid_tag.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_grid_tag_faces.argtypes = [C.c_void_p, C.c_int, _i64p, _i64p]
        L.ref_grid_wls_r.restype = C.c_int
        L.ref_grid_wls_r.argtypes = [C.c_void_p, _dp]
        L.ref_compute_dudt.restype = C.c_int
        L.ref_compute_dudt.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        if hasattr(L, "ref_compute_dudt_moving"):
            L.ref_compute_dudt_moving.restype = C.c_int
            L.ref_compute_dudt_moving.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
        L.ref_time_dudt.restype = C.c_double
        L.ref_time_dudt.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int]
        L.ref_intermediates.argtypes = [C.c_void_p] + [C.c_void_p] * 15
        L.ref_estimate_dt.restype = C.c_double
        L.ref_estimate_dt.argtypes = [C.c_void_p, _dp]
        L.ref_count_bad_cells.restype = C.c_int64
        L.ref_count_bad_cells.argtypes = [C.c_void_p, _dp]
        L.ref_prim_to_cons.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_cons_to_prim.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_l2_norms.argtypes = [C.c_void_p, _dp]
        L.ref_solve.restype = C.c_int
        L.ref_solve.argtypes = [C.c_char_p]
        L.ref_gmres_dense.restype = C.c_int
        L.ref_gmres_dense.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, _dp,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_init(int(num_threads))
        self.num_threads = L.ref_num_threads()

    def solve(self, case_dir):
        return self.lib.ref_solve(os.fsencode(case_dir))

    def gmres_dense(self, A, b, x0, max_iters, tol):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64)
        n_it = C.c_int(0)
        res = C.c_double(0)
        ok = self.lib.ref_gmres_dense(A, b, len(b), max_iters, tol, x, C.byref(n_it), C.byref(res))
        return x, bool(ok), n_it.value, res.value


class RefCase:
    """One reference GridBlock + FiniteVolume (real or dual) built from an SU2
    file and a config dict with the reference's config.json schema."""

    def __init__(self, su2_path, config, dual=False, num_threads=0):
        self.rl = RefLib.get(num_threads)
        self.lib = self.rl.lib
        self.dual = bool(dual)
        self.h = self.lib.ref_case_create(os.fsencode(su2_path), json.dumps(config).encode(),
                                          int(self.dual))
        if not self.h:
            raise RuntimeError("reference failed to build the case")
        s = np.zeros(8, dtype=np.int64)
        self.lib.ref_case_sizes(self.h, s)
        (self.dim, self.n_cells, self.n_ghost, self.n_faces, self.n_vertices,
         self.n_cf, self.n_tags, self.n_cons) = [int(v) for v in s]
        self.n_total = self.n_cells + self.n_ghost

    def close(self):
        if self.h:
            self.lib.ref_case_destroy(self.h)
            self.h = None

    # ---- grid ----
    def grid_arrays(self):
        """The reference GridBlock's arrays, reference numbering (dict of numpy arrays)."""
        nf, n, nt = self.n_faces, self.n_cells, self.n_total
        g = dict(dim=self.dim, n_cells=n, n_ghost=self.n_ghost, n_faces=nf)
        g["face_left"] = np.zeros(nf, np.int64)
        g["face_right"] = np.zeros(nf, np.int64)
        g["face_area"] = np.zeros(nf)
        for k in ("face_norm", "face_tan1", "face_tan2", "face_centre"):
            g[k] = np.zeros((nf, 3))
        self.lib.ref_grid_faces(self.h, g["face_left"], g["face_right"], g["face_area"],
                                g["face_norm"], g["face_tan1"], g["face_tan2"], g["face_centre"])
        g["cell_face_offsets"] = np.zeros(n + 1, np.int64)
        g["cell_face_ids"] = np.zeros(self.n_cf, np.int64)
        g["cell_outsigns"] = np.zeros(self.n_cf, np.int32)
        g["cell_neighbours"] = np.zeros(self.n_cf, np.int64)
        g["cell_volume"] = np.zeros(n)
        g["cell_centroid"] = np.zeros((nt, 3))
        self.lib.ref_grid_cells(self.h, g["cell_face_offsets"], g["cell_face_ids"],
                                g["cell_outsigns"], g["cell_neighbours"], g["cell_volume"],
                                g["cell_centroid"])
        tags = {}
        buf = C.create_string_buffer(256)
        for i in range(self.n_tags):
            m = self.lib.ref_grid_tag(self.h, i, buf, 256)
            faces = np.zeros(m, np.int64)
            ghosts = np.zeros(m, np.int64)
            self.lib.ref_grid_tag_faces(self.h, i, faces, ghosts)
            tags[buf.value.decode()] = dict(faces=faces, ghosts=ghosts)
        g["tags"] = tags
        return g

    def wls_r(self):
        out = np.zeros((self.n_cells, 6))
        nc = self.lib.ref_grid_wls_r(self.h, out)
        return out.reshape(-1)[: self.n_cells * nc].reshape(self.n_cells, nc) if nc else None

    # ---- hot path ----
    def compute_dudt(self, prim, prim_dual=None, allow_reconstruction=True):
        """prim: (n_total, 7) rho,p,T,e,vx,vy,vz. Returns (dudt, prim_after[, dudt_du, prim_du])."""
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        dr = np.zeros((self.n_cells, self.n_cons))
        if self.dual:
            pd = (np.zeros_like(pr) if prim_dual is None
                  else np.array(prim_dual, dtype=np.float64, order="C").reshape(self.n_total, 7))
            dd = np.zeros_like(dr)
            self.lib.ref_compute_dudt(self.h, _ptr(pr), _ptr(pd), _ptr(dr), _ptr(dd),
                                      int(allow_reconstruction))
            return dr, pr, dd, pd
        self.lib.ref_compute_dudt(self.h, _ptr(pr), None, _ptr(dr), None, int(allow_reconstruction))
        return dr, pr

    def compute_dudt_moving(self, prim, prim_dual=None, allow_reconstruction=True):
        """the moving-grid overload (config.grid.motion enabled). Returns (dudt, prim_after, face_vel) or, dual,
        (dudt, prim_after, dudt_du, prim_du, face_vel)."""
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        dr = np.zeros((self.n_cells, self.n_cons))
        fv = np.zeros((self.n_faces, 3))
        if self.dual:
            pd = (np.zeros_like(pr) if prim_dual is None
                  else np.array(prim_dual, dtype=np.float64, order="C").reshape(self.n_total, 7))
            dd = np.zeros_like(dr)
            rc = self.lib.ref_compute_dudt_moving(self.h, _ptr(pr), _ptr(pd), _ptr(dr), _ptr(dd),
                                                  int(allow_reconstruction), _ptr(fv))
            assert rc == 0, "the reference grid is not moving"
            return dr, pr, dd, pd, fv
        rc = self.lib.ref_compute_dudt_moving(self.h, _ptr(pr), None, _ptr(dr), None, int(allow_reconstruction), _ptr(fv))
        assert rc == 0, "the reference grid is not moving"
        return dr, pr, fv

    def time_dudt(self, prim, nrep, allow_reconstruction=True):
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        return self.lib.ref_time_dudt(self.h, pr, int(nrep), int(allow_reconstruction))

    def intermediates(self):
        n, nf = self.n_cells, self.n_faces
        names3 = ["grad_rho", "grad_u", "grad_p", "grad_T", "grad_vx", "grad_vy", "grad_vz"]
        names1 = ["lim_rho", "lim_u", "lim_p", "lim_T", "lim_vx", "lim_vy", "lim_vz"]
        out = {k: np.full((n, 3), np.nan) for k in names3}
        out.update({k: np.full(n, np.nan) for k in names1})
        out["flux"] = np.zeros((nf, self.n_cons))
        args = [_ptr(out[k]) for k in names3 + names1] + [_ptr(out["flux"])]
        self.lib.ref_intermediates(self.h, *args)
        return out

    def estimate_dt(self, prim):
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        return self.lib.ref_estimate_dt(self.h, pr)

    def count_bad_cells(self, prim):
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        return int(self.lib.ref_count_bad_cells(self.h, pr))

    def prim_to_cons(self, prim, prim_dual=None):
        pr = np.array(prim, dtype=np.float64, order="C").reshape(self.n_total, 7)
        cr = np.zeros((self.n_total, self.n_cons))
        if self.dual:
            pd = np.zeros_like(pr) if prim_dual is None else np.ascontiguousarray(prim_dual)
            cd = np.zeros_like(cr)
            self.lib.ref_prim_to_cons(self.h, _ptr(pr), _ptr(pd), _ptr(cr), _ptr(cd))
            return cr, cd
        self.lib.ref_prim_to_cons(self.h, _ptr(pr), None, _ptr(cr), None)
        return cr

    def cons_to_prim(self, cons, cons_dual=None):
        cr = np.array(cons, dtype=np.float64, order="C").reshape(self.n_total, self.n_cons)
        pr = np.zeros((self.n_total, 7))
        if self.dual:
            cd = np.zeros_like(cr) if cons_dual is None else np.ascontiguousarray(cons_dual)
            pd = np.zeros_like(pr)
            self.lib.ref_cons_to_prim(self.h, _ptr(cr), _ptr(cd), _ptr(pr), _ptr(pd))
            return pr, pd
        self.lib.ref_cons_to_prim(self.h, _ptr(cr), None, _ptr(pr), None)
        return pr

    def l2_norms(self):
        out = np.zeros(6)
        self.lib.ref_l2_norms(self.h, out)
        return out


def run_prep(case_dir, job="job.py"):
    """Run the reference's own prep.py (src/ibis/ibis/commands/prep/prep.py:1178)
    in ``case_dir`` (which must contain job.py and its grid) in a subprocess.
    Produces config/config.json, io/grid/0000/block_0000.su2, io/flow/0000/*."""
    cmds = os.path.join(REFERENCE_ROOT, "src/ibis/ibis/commands")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(HERE, "_ref"), os.path.join(cmds, "prep"), cmds, env.get("PYTHONPATH", "")])
    os.makedirs(os.path.join(case_dir, "log"), exist_ok=True)
    code = f"import prep; prep.main({job!r}, {os.path.join(REFERENCE_ROOT, 'share')!r})"
    subprocess.run([sys.executable, "-c", code], cwd=case_dir, env=env, check=True)
    with open(os.path.join(case_dir, "config", "config.json")) as f:
        return json.load(f)


def read_flow(case_dir, index, dim):
    """Native binary flow snapshot (io/native.cpp:186-259): raw f64 per valid cell."""
    d = os.path.join(case_dir, "io", "flow", f"{index:04d}")
    names = ["T", "p", "vx", "vy"] + (["vz"] if dim == 3 else [])
    return {k: np.fromfile(os.path.join(d, k)) for k in names}


def copy_case(src, dst):
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    return dst
