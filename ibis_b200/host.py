"""ctypes binding of libibis_b200_host.so (include/ibis_b200_host.h): the C++ host solvers
(Solver::solve, RungeKutta, SteadyState, Jfnk, Gmres, FGmres, CflSchedule) above the C ABI.

``solve(ctx, config, prim0)`` is ``make_solver(config, ...)->solve()`` of the reference
(/root/reference/src/solvers/solvers/solver.cpp:18-84) with ``config`` the dict of
``config/config.json`` and the initial flow snapshot passed as an array.
"""
import ctypes as C
import os

import numpy as np

from .lib import IbisB200Error, Library

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_HOST_LIB = os.path.join(HERE, "lib", "libibis_b200_host.so")
_dp = C.POINTER(C.c_double)


class CflConfig(C.Structure):
    _fields_ = [("type", C.c_int32), ("n", C.c_int32), ("value", C.c_double), ("times", _dp),
                ("cfls", _dp), ("growth_threshold", C.c_double), ("power", C.c_double),
                ("start_cfl", C.c_double), ("max_cfl", C.c_double)]


class RkConfig(C.Structure):
    _fields_ = [("n_stages", C.c_int32), ("fused_step", C.c_int32), ("a", _dp), ("b", _dp),
                ("cfl", CflConfig), ("max_time", C.c_double), ("max_step", C.c_int64),
                ("dt_init", C.c_double), ("print_frequency", C.c_int32),
                ("plot_every_n_steps", C.c_int32), ("plot_frequency", C.c_double),
                ("residuals_every_n_steps", C.c_int32), ("distinct_stages", C.c_int32),
                ("residual_frequency", C.c_double), ("moving_grid", C.c_int32), ("reserved", C.c_int32)]


class LinearSolverConfig(C.Structure):
    _fields_ = [("type", C.c_int32), ("max_iters", C.c_int32), ("tol", C.c_double),
                ("max_preconditioner_iters", C.c_int32), ("reserved", C.c_int32),
                ("preconditioner_tolerance", C.c_double)]


class SteadyConfig(C.Structure):
    _fields_ = [("cfl", CflConfig), ("linear_solver", LinearSolverConfig), ("max_steps", C.c_int64),
                ("tolerance", C.c_double), ("print_frequency", C.c_int32), ("plot_frequency", C.c_int32),
                ("diagnostics_frequency", C.c_int32), ("reserved", C.c_int32)]


class SolveReport(C.Structure):
    _fields_ = [("status", C.c_int32), ("stop_reason", C.c_int32), ("steps", C.c_int64),
                ("bad_cells", C.c_int64), ("t", C.c_double), ("dt", C.c_double),
                ("stable_dt", C.c_double), ("residual_norms", C.c_double * 6),
                ("relative_residual", C.c_double), ("linear_iters", C.c_int64),
                ("last_linear_success", C.c_int32), ("last_linear_iters", C.c_int32),
                ("last_linear_residual", C.c_double), ("wall_seconds", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "residual_norms"}
        d["residual_norms"] = list(self.residual_norms)
        return d


PLOT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_double, C.c_int)
RESIDUAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_double, C.c_double, _dp)


class Observers(C.Structure):
    _fields_ = [("user", C.c_void_p), ("plot", PLOT_FN), ("residuals", RESIDUAL_FN)]


def _cfl(c, keep):
    """solver.cfl of config.json -> CflConfig (solvers/cfl.cpp:72-84)."""
    out = CflConfig()
    t = c["type"]
    if t == "constant":
        out.type, out.value = 0, float(c["value"])
    elif t == "linear_interpolate":
        times = np.ascontiguousarray(c["times"], dtype=np.float64)
        cfls = np.ascontiguousarray(c["cfls"], dtype=np.float64)
        keep += [times, cfls]
        out.type, out.n = 1, len(times)
        out.times = times.ctypes.data_as(_dp)
        out.cfls = cfls.ctypes.data_as(_dp)
    elif t == "residual_based":
        out.type = 2
        out.growth_threshold = float(c["growth_threshold"])
        out.power = float(c["power"])
        out.start_cfl = float(c["start_cfl"])
        out.max_cfl = float(c["max_cfl"])
    else:
        raise IbisB200Error("Unknown CFL schedule")
    return out


def _linear_solver(c):
    """solver.linear_solver -> LinearSolverConfig (linear_algebra/gmres.cpp:116-117,228-232,305-318)."""
    out = LinearSolverConfig()
    if c["type"] == "gmres":
        out.type, out.max_iters, out.tol = 0, int(c["max_iters"]), float(c["tol"])
    elif c["type"] == "fgmres":
        out.type, out.max_iters, out.tol = 1, int(c["max_iters"]), float(c["tolerance"])
        out.max_preconditioner_iters = int(c["max_preconditioner_iters"])
        out.preconditioner_tolerance = float(c["preconditioner_tolerance"])
    else:
        raise IbisB200Error("Unknown linear solver")
    return out


class HostLibrary:
    _cache = {}

    @classmethod
    def load(cls, path=None, core_path=None):
        path = os.path.abspath(path or os.environ.get("IBIS_B200_HOST_LIB", DEFAULT_HOST_LIB))
        if path not in cls._cache:
            cls._cache[path] = cls(path, core_path)
        return cls._cache[path]

    def __init__(self, path, core_path=None):
        Library.load(core_path)          # the core library must be in the process first
        if not os.path.exists(path):
            raise IbisB200Error(f"{path} not found: run __graft_entry__.build()")
        L = self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        L.ibis_b200_host_run_rk.argtypes = [vp, C.POINTER(RkConfig), vp, vp, C.POINTER(SolveReport),
                                            C.POINTER(Observers)]
        L.ibis_b200_host_run_steady.argtypes = [vp, C.POINTER(SteadyConfig), vp, vp, C.POINTER(SolveReport),
                                                C.POINTER(Observers)]
        L.ibis_b200_host_gmres_dense.argtypes = [vp, C.POINTER(LinearSolverConfig), vp, vp, C.c_int, vp,
                                                 C.POINTER(C.c_int), _dp]
        for f in (L.ibis_b200_host_run_rk, L.ibis_b200_host_run_steady, L.ibis_b200_host_gmres_dense):
            f.restype = C.c_int
        L.ibis_b200_host_last_error.restype = C.c_char_p


def solve(ctx, config, prim0, fused_step=True, on_plot=None, on_residuals=None, host_lib=None,
          distinct_stages=False):
    """Run the solver named by config["solver"] on context ``ctx`` from the primitive state
    ``prim0`` (n_total x 7).  Returns (final primitive state, report dict)."""
    H = HostLibrary.load(host_lib, ctx.L.path)
    sc = config["solver"]
    keep = []
    prim0 = np.ascontiguousarray(prim0, dtype=np.float64).reshape(ctx.n_total, 7)
    out = np.zeros_like(prim0)
    rep = SolveReport()
    obs = Observers()

    def _plot(_user, step, t, fs):
        if on_plot:
            on_plot(int(step), float(t), ctx.fs_download(fs) if not _is_dual(sc) else ctx.fs_download(fs, dual=True)[0])

    def _res(_user, step, t, wall, norms):
        if on_residuals:
            on_residuals(int(step), float(t), [norms[i] for i in range(6)])

    plot_cb, res_cb = PLOT_FN(_plot), RESIDUAL_FN(_res)
    obs.plot, obs.residuals = plot_cb, res_cb
    if sc["name"] == "runge_kutta":
        tab = sc["butcher_tableau"]
        b = np.ascontiguousarray(tab["b"], dtype=np.float64)
        n = len(b)
        a = np.array([tab["a"][i - 1][j] for i in range(1, n) for j in range(i)] or [0.0], dtype=np.float64)
        cfg = RkConfig()
        cfg.n_stages, cfg.fused_step = n, int(bool(fused_step))
        cfg.distinct_stages = int(bool(distinct_stages))
        cfg.a, cfg.b = a.ctypes.data_as(_dp), b.ctypes.data_as(_dp)
        cfg.cfl = _cfl(sc["cfl"], keep)
        cfg.max_time, cfg.max_step = float(sc["max_time"]), int(sc["max_step"])
        cfg.dt_init = float(sc["dt_init"])
        cfg.print_frequency = int(sc["print_frequency"])
        cfg.plot_every_n_steps = int(sc["plot_every_n_steps"])
        cfg.plot_frequency = float(sc["plot_frequency"])
        cfg.residuals_every_n_steps = int(sc["residuals_every_n_steps"])
        cfg.residual_frequency = float(sc["residual_frequency"])
        cfg.moving_grid = int(bool(config["grid"].get("motion", {}).get("enabled", False)))
        if cfg.moving_grid:
            cfg.fused_step = 0
        rc = H.lib.ibis_b200_host_run_rk(ctx.h, C.byref(cfg), prim0.ctypes.data, out.ctypes.data,
                                         C.byref(rep), C.byref(obs))
    elif sc["name"] == "steady_state":
        cfg = SteadyConfig()
        cfg.cfl = _cfl(sc["cfl"], keep)
        cfg.linear_solver = _linear_solver(sc["linear_solver"])
        cfg.max_steps, cfg.tolerance = int(sc["max_steps"]), float(sc["tolerance"])
        cfg.print_frequency = int(sc["print_frequency"])
        cfg.plot_frequency = int(sc["plot_frequency"])
        cfg.diagnostics_frequency = int(sc["diagnostics_frequency"])
        rc = H.lib.ibis_b200_host_run_steady(ctx.h, C.byref(cfg), prim0.ctypes.data, out.ctypes.data,
                                             C.byref(rep), C.byref(obs))
    else:
        raise IbisB200Error("Unknown solver " + str(sc["name"]))     # solver.cpp:80-82
    if rc < 0:
        raise IbisB200Error(f"solve failed ({rc}): {H.lib.ibis_b200_host_last_error().decode()}")
    return out, rep.as_dict()


def _is_dual(solver_config):
    return solver_config["name"] == "steady_state"


def gmres_dense(ctx, A, b, x0, linear_solver, host_lib=None):
    """Gmres / FGmres of the host library on a dense system (known-answer tests)."""
    H = HostLibrary.load(host_lib, ctx.L.path)
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    cfg = _linear_solver(linear_solver)
    n_it, res = C.c_int(0), C.c_double(0)
    rc = H.lib.ibis_b200_host_gmres_dense(ctx.h, C.byref(cfg), A.ctypes.data, b.ctypes.data, len(b),
                                          x.ctypes.data, C.byref(n_it), C.byref(res))
    if rc < 0:
        raise IbisB200Error(f"gmres_dense failed ({rc}): {H.lib.ibis_b200_host_last_error().decode()}")
    return x, bool(rc), n_it.value, res.value
