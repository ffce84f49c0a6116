"""Whole-solver parity cases: mesh + config (incl. the "solver" object) + initial snapshot.

tests/golden/make_solve_golden.py runs the REAL reference (make_solver(config)->solve(),
/root/reference/src/solvers/solvers/solver.cpp:18-84) on each of them and stores the final
flow snapshot; the tests run the C++ host solvers of ibis_b200/host above the C ABI on the
same inputs and compare the flow fields (north-star tolerance 1e-9)."""
import numpy as np

import cases
from ibis_b200 import config as cfg
from ibis_b200 import grid as gridmod
from ibis_b200 import su2

R, CV = 287.0, 717.5


def snapshot_state(g, T, p, vx, vy, vz):
    """FVIO::read (io/native.cpp:264-339): T, p, v of the valid cells from the snapshot,
    rho and e from update_thermo_from_pT (gas/gas_model.h:158-170); ghost rows stay zero."""
    n, nt = g.n_cells, g.n_total
    prim = np.zeros((nt, 7))
    prim[:n, 2], prim[:n, 1] = T, p
    prim[:n, 0] = p / (R * T)
    prim[:n, 3] = CV * T
    prim[:n, 4], prim[:n, 5], prim[:n, 6] = vx, vy, vz
    return prim


def _uniform(g, fs, dim):
    n = g.n_cells
    o = np.ones(n)
    return fs["T"] * o, fs["p"] * o, fs["vx"] * o, fs["vy"] * o, (fs["vz"] if dim == 3 else 0.0) * o


def _wavy(g, fs, dim, amp=0.02):
    """Free stream with a smooth +-amp perturbation.  (An exactly uniform field makes the dual
    residual ill-conditioned: the Barth-Jespersen limiter's derivative is ~1/epsilon = 1e25 where
    the gradient vanishes, and the steady solver then depends on the summation ORDER of its dot
    products -- the reference itself differs by 1.5 K between 1 and 8 OpenMP threads there.)"""
    n = g.n_cells
    x, y = g.cell_centroid[:n, 0], g.cell_centroid[:n, 1]
    w = 1.0 + amp * np.sin(3.0 * x + 1.0) * np.cos(2.0 * y)
    return fs["T"] * w, fs["p"] * (2.0 - w), fs["vx"] * w, fs["vy"] + 0.0 * w + amp * fs["vx"] * np.sin(2.0 * x), \
        (fs["vz"] if dim == 3 else 0.0) * w


def _perturbed(g, dim, mach):
    full = cases.perturbed_state(g, dim, mach)
    n = g.n_cells
    return full[:n, 2], full[:n, 1], full[:n, 4], full[:n, 5], full[:n, 6]


FS_M3 = cfg.flow_state(p=101325.0, T=300.0, vx=3.0 * np.sqrt(1.4 * 287.0 * 300.0))
FS_PLATE = cfg.flow_state(p=1.013e3, T=300.0, vx=1390.0)
FS_SUB = cfg.flow_state(rho=1.225, T=300.0, vx=200.0, vy=20.0)

# name -> (mesh factory, bcs, initial-field factory (g, dim) -> T,p,vx,vy,vz, make_config kwargs)
SOLVE_CASES = {
    # examples/wedge/2D/transient: Mach 3 inviscid ramp, explicit
    "rk_wedge2d": (lambda: su2.wedge_2d(24, 16),
                   {"inflow": cfg.supersonic_inflow(FS_M3), "outflow": cfg.supersonic_outflow(),
                    "wall": cfg.slip_wall()},
                   lambda g, d: _uniform(g, FS_M3, d),
                   dict(flux_calculator="ausmdv",
                        solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5, max_step=40, max_time=1.0))),
    # examples/forward_facing_step: Mach 3, MUSCL + limiter, shock dominated
    "rk_ffs2d": (lambda: su2.forward_facing_step_2d(1.0 / 10),
                 {"inflow": cfg.supersonic_inflow(FS_M3), "outflow": cfg.supersonic_outflow(),
                  "wall": cfg.slip_wall()},
                 lambda g, d: _uniform(g, FS_M3, d),
                 dict(flux_calculator="hanel",
                      solver=cfg.runge_kutta(method="rk4", cfl=0.4, max_step=30, max_time=1.0, dt_init=1e-7))),
    # examples/flat_plate: viscous laminar boundary layer
    "rk_plate2d_viscous": (lambda: su2.flat_plate_2d(24, 16, height=0.05),
                           {"inflow": cfg.supersonic_inflow(FS_PLATE), "outflow": cfg.supersonic_outflow(),
                            "wall": cfg.fixed_temperature_no_slip_wall(300.0)},
                           lambda g, d: _uniform(g, FS_PLATE, d),
                           dict(flux_calculator="ausmdv", viscous=True,
                                solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5, max_step=30, max_time=1.0))),
    # examples/naca0012: viscous, subsonic in/outflow
    "rk_naca_viscous": (lambda: su2.naca0012_omesh(32, 12),
                        {"wall": cfg.adiabatic_no_slip_wall(), "inflow": cfg.subsonic_inflow(FS_SUB),
                         "outflow": cfg.subsonic_outflow(FS_SUB["p"])},
                        lambda g, d: _uniform(g, FS_SUB, d),
                        dict(flux_calculator="ausmdv", viscous=True,
                             solver=cfg.runge_kutta(method="midpoint", cfl=0.4, max_step=25, max_time=1.0))),
    # 3-D box, ldfss2, linear CFL schedule, stops on max_time
    "rk_box3d": (lambda: su2.box_hex(6, 5, 4, jitter=0.2), cases.box3_bcs(),
                 lambda g, d: _perturbed(g, d, 1.0),
                 dict(flux_calculator="ldfss2",
                      solver=cfg.runge_kutta(method="euler", max_step=200, max_time=2.0e-4,
                                             cfl={"type": "linear_interpolate", "times": [0.0, 1e-4, 1.0],
                                                  "cfls": [0.1, 0.4, 0.4]}))),
    # examples/wedge/2D/steady_state: JFNK + GMRES
    "steady_wedge2d": (lambda: su2.wedge_2d(16, 10),
                       {"inflow": cfg.supersonic_inflow(FS_M3), "outflow": cfg.supersonic_outflow(),
                        "wall": cfg.slip_wall()},
                       lambda g, d: _wavy(g, FS_M3, d),
                       dict(flux_calculator="ausmdv",
                            solver=cfg.steady_state(cfl={"type": "residual_based", "growth_threshold": 0.9,
                                                         "power": 1.0, "start_cfl": 0.5, "max_cfl": 1e4},
                                                    max_steps=12, tolerance=1e-10,
                                                    linear_solver={"type": "gmres", "max_iters": 30,
                                                                   "tol": 1e-2}))),
    # tests/minimal/steady_state_uniform_stream flavour on a perturbed 2-D field, constant CFL
    "steady_quad2d": (lambda: su2.box_quad(8, 6, jitter=0.15), cases.quad2_bcs(),
                      lambda g, d: _perturbed(g, d, 1.0),
                      dict(flux_calculator="hanel",
                           solver=cfg.steady_state(cfl=5.0, max_steps=8, tolerance=1e-12,
                                                   linear_solver={"type": "gmres", "max_iters": 40,
                                                                  "tol": 1e-3}))),
    # flat_plate steady flavour: FGMRES with the first-order preconditioner, viscous
    "steady_plate_fgmres": (lambda: su2.flat_plate_2d(12, 8, height=0.05),
                            {"inflow": cfg.supersonic_inflow(FS_PLATE), "outflow": cfg.supersonic_outflow(),
                             "wall": cfg.adiabatic_no_slip_wall()},
                            lambda g, d: _uniform(g, FS_PLATE, d),
                            dict(flux_calculator="ausmdv", viscous=True,
                                 solver=cfg.steady_state(cfl=2.0, max_steps=5, tolerance=1e-12,
                                                         linear_solver={"type": "fgmres", "max_iters": 10,
                                                                        "tolerance": 1e-2,
                                                                        "max_preconditioner_iters": 60,
                                                                        "preconditioner_tolerance": 1e-1}))),
    # examples/flat_plate/2D/truncated_plate: boundary-layer inflow profile (cubic splines in the ghost
    # centroid's y), viscous, explicit.  Knots: the domain top (0.05) lies below the third-last knot, so no
    # ghost centroid falls in the last two spline intervals (reference out-of-bounds reads)
    "rk_plate_viscous_bl_inflow": (lambda: su2.flat_plate_2d(24, 16, height=0.05),
                           {"inflow": cfg.boundary_layer_inflow(
                               [0.0, 0.002, 0.004, 0.007, 0.011, 0.016, 0.03, 0.051, 0.06, 0.07],
                               [0.0, 420.0, 760.0, 1050.0, 1250.0, 1350.0, 1388.0, 1390.0, 1390.0, 1390.0],
                               [300.0, 480.0, 520.0, 470.0, 390.0, 330.0, 302.0, 300.0, 300.0, 300.0], 1.013e3),
                            "outflow": cfg.supersonic_outflow(),
                            "wall": cfg.fixed_temperature_no_slip_wall(300.0)},
                           lambda g, d: _uniform(g, FS_PLATE, d),
                           dict(flux_calculator="ausmdv", viscous=True,
                                solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5, max_step=30, max_time=1.0))),
    # constant_flux inflow (prep.py bow_shock_fit = constant_flux, here on a static grid), explicit
    "rk_quad2d_cflux": (lambda: su2.box_quad(10, 8, jitter=0.15),
                        dict(cases.quad2_bcs(), left=cfg.constant_flux(cases.FS2)),
                        lambda g, d: _perturbed(g, d, 1.0),
                        dict(flux_calculator="ausmdv",
                             solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5, max_step=25, max_time=1.0))),
    # ... and through the dual residual of the steady solver
    "steady_quad2d_cflux": (lambda: su2.box_quad(8, 6, jitter=0.15),
                            dict(cases.quad2_bcs(), left=cfg.constant_flux(cases.FS2)),
                            lambda g, d: _perturbed(g, d, 1.0),
                            dict(flux_calculator="hanel",
                                 solver=cfg.steady_state(cfl=5.0, max_steps=8, tolerance=1e-12,
                                                         linear_solver={"type": "gmres", "max_iters": 40,
                                                                        "tol": 1e-3}))),
    # moving grids: rigid_body_translation driver, vertices displaced and the geometry recomputed every stage
    # (runge_kutta.cpp:136,157-196)
    "rk_quad2d_moving": (lambda: su2.box_quad(10, 8, jitter=0.15), cases.quad2_bcs(),
                         lambda g, d: _perturbed(g, d, 1.0),
                         dict(flux_calculator="ausmdv", motion=cfg.rigid_body_translation(150.0, -60.0, 0.0),
                              solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5, max_step=20, max_time=1.0))),
    "rk_box3d_moving": (lambda: su2.box_hex(5, 4, 4, jitter=0.2), cases.box3_bcs(),
                        lambda g, d: _perturbed(g, d, 1.0),
                        dict(flux_calculator="hanel", motion=cfg.rigid_body_translation(-80.0, 40.0, 25.0),
                             solver=cfg.runge_kutta(method="midpoint", cfl=0.4, max_step=15, max_time=1.0))),
    # 3-D steady
    "steady_box3d": (lambda: su2.box_hex(4, 4, 3, jitter=0.1), cases.box3_wall_bcs(),
                     lambda g, d: _perturbed(g, d, 1.0),
                     dict(flux_calculator="ausmdv",
                          solver=cfg.steady_state(cfl=3.0, max_steps=6, tolerance=1e-12,
                                                  linear_solver={"type": "gmres", "max_iters": 40,
                                                                 "tol": 1e-3}))),
}


def build_solve_case(name):
    """-> (mesh, config, Grid, prim0 (n_total x 7), snapshot dict T,p,vx,vy,vz)"""
    mesh_f, bcs, init_f, kw = SOLVE_CASES[name]
    mesh = mesh_f()
    config = cfg.make_config(mesh.dim, bcs, **kw)
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    T, p, vx, vy, vz = init_f(g, mesh.dim)
    prim0 = snapshot_state(g, T, p, vx, vy, vz)
    return mesh, config, g, prim0, {"T": T, "p": p, "vx": vx, "vy": vy, "vz": vz}
