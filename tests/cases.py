"""Shared parity cases: mesh + boundary conditions + configuration + seeded input state.

Used by tests/golden/make_golden.py (which runs the real reference on them), by the CPU
tests (host emulation of the kernel bodies + the C-ABI layer against the golden vectors) and by
the GPU parity tests (CUDA path through the C ABI against the golden vectors and, when it
travelled to the box, against the compiled reference oracle/_ref itself)."""
import numpy as np

from ibis_b200 import config as cfg
from ibis_b200 import grid as gridmod
from ibis_b200 import su2

R, CV = 287.0, 717.5


def perturbed_state(g, dim, mach_scale=1.0, seed=1234):
    """SURVEY.md section 8(d) field (ii): smooth perturbation + U(-1,1) m/s noise by cell index.
    One deviation from the survey's formula, kept because every golden vector and both bench arms were generated
    with it: the density perturbation ends in cos(2 pi z) where the survey writes sin(2 pi z).  Same amplitude,
    same smoothness, same field on the GPU and the reference arm: no effect on parity or on the comparison."""
    n = g.n_total
    x, y, z = g.cell_centroid[:, 0], g.cell_centroid[:, 1], g.cell_centroid[:, 2]
    rng = np.random.default_rng(seed)
    rho = 1.225 * (1 + 0.2 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) * np.cos(2 * np.pi * z))
    T = 300 * (1 + 0.2 * np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y))
    s = 1 + 0.1 * np.sin(2 * np.pi * (x + y + z))
    vx = mach_scale * 1000 * s + rng.uniform(-1, 1, n)
    vy = mach_scale * 500 * s + rng.uniform(-1, 1, n)
    vz = (mach_scale * 400 * s + rng.uniform(-1, 1, n)) if dim == 3 else np.zeros(n)
    p = rho * R * T
    e = CV * T
    return np.ascontiguousarray(np.stack([rho, p, T, e, vx, vy, vz], axis=1))


def uniform_state(g, dim, fs):
    n = g.n_total
    row = [fs["rho"], fs["p"], fs["T"], fs["energy"], fs["vx"], fs["vy"], fs["vz"] if dim == 3 else 0.0]
    return np.ascontiguousarray(np.tile(np.array(row), (n, 1)))


FS3 = cfg.flow_state(rho=1.225, T=300.0, vx=1000, vy=500, vz=400)
FS2 = cfg.flow_state(rho=1.225, T=300.0, vx=1000, vy=500)
FS2_SUB = cfg.flow_state(rho=1.225, T=300.0, vx=200, vy=100)


def box3_bcs():
    return {"west": cfg.supersonic_inflow(FS3), "south": cfg.supersonic_inflow(FS3),
            "bottom": cfg.supersonic_inflow(FS3), "east": cfg.supersonic_outflow(),
            "north": cfg.supersonic_outflow(), "top": cfg.supersonic_outflow()}


def box3_wall_bcs():
    b = box3_bcs()
    b["bottom"] = cfg.slip_wall()
    b["top"] = cfg.slip_wall()
    return b


def box3_visc_bcs():
    b = box3_bcs()
    b["bottom"] = cfg.adiabatic_no_slip_wall()
    b["top"] = cfg.fixed_temperature_no_slip_wall(400.0)
    return b


def quad2_bcs():
    return {"left": cfg.supersonic_inflow(FS2), "bottom": cfg.slip_wall(),
            "right": cfg.supersonic_outflow(), "top": cfg.supersonic_outflow()}


def quad2_sub_bcs():
    return {"left": cfg.subsonic_inflow(FS2_SUB), "bottom": cfg.slip_wall(),
            "right": cfg.subsonic_outflow(101325.0), "top": cfg.subsonic_outflow(101325.0)}


def quad2_visc_bcs():
    return {"left": cfg.supersonic_inflow(FS2), "bottom": cfg.fixed_temperature_no_slip_wall(500.0),
            "right": cfg.supersonic_outflow(), "top": cfg.adiabatic_no_slip_wall()}


def wedge_bcs():
    return {"inflow": cfg.supersonic_inflow(FS2_M3), "outflow": cfg.supersonic_outflow(),
            "wall": cfg.slip_wall()}


FS2_M3 = cfg.flow_state(p=101325.0, T=300.0, vx=3.0 * np.sqrt(1.4 * 287.0 * 300.0))

# name -> (mesh factory, bcs factory, mach_scale, config kwargs)
CASES = {}


def _add(name, mesh, bcs, mach=1.0, **kw):
    CASES[name] = (mesh, bcs, mach, kw)


_hex = lambda: su2.box_hex(8, 7, 6, jitter=0.2)            # noqa: E731
_quad = lambda: su2.box_quad(12, 9, jitter=0.2)            # noqa: E731
for _fc in ("ausmdv", "hanel", "ldfss2", "rusanov"):
    _add("box3d_" + _fc, _hex, box3_bcs, flux_calculator=_fc)
    _add("quad2d_" + _fc, _quad, quad2_bcs, flux_calculator=_fc)
_add("box3d_order1", _hex, box3_bcs, flux_calculator="ausmdv", reconstruction_order=1)
_add("box3d_unlimited", _hex, box3_bcs, flux_calculator="ausmdv", limiter="unlimited")
_add("box3d_rho_p", _hex, box3_bcs, flux_calculator="ausmdv", thermo_interpolator="rho_p")
_add("box3d_rho_T", _hex, box3_bcs, flux_calculator="ausmdv", thermo_interpolator="rho_T")
_add("box3d_subsonic", _hex, box3_bcs, 0.2, flux_calculator="ausmdv")
_add("box3d_regular", lambda: su2.box_hex(6), box3_bcs, flux_calculator="ausmdv")
_add("mixed3d_walls", lambda: su2.mixed_3d(4), box3_wall_bcs, 0.3, flux_calculator="ausmdv")
_add("tet3d", lambda: su2.box_tet(4), box3_wall_bcs, 0.3, flux_calculator="hanel")
_add("tri2d", lambda: su2.box_quad(12, 9, jitter=0.2, triangulate=True), quad2_bcs, 0.4,
     flux_calculator="ausmdv")
_add("quad2d_subsonic_bcs", _quad, quad2_sub_bcs, 0.2, flux_calculator="ausmdv")
_add("quad2d_viscous", _quad, quad2_visc_bcs, 0.5, flux_calculator="ausmdv", viscous=True)
_add("quad2d_viscous_o1", _quad, quad2_visc_bcs, 0.5, flux_calculator="hanel", viscous=True,
     reconstruction_order=1)
_add("box3d_viscous", lambda: su2.box_hex(6, 5, 4, jitter=0.2), box3_visc_bcs, 0.5,
     flux_calculator="ausmdv", viscous=True)
_add("wedge2d", lambda: su2.wedge_2d(24, 16), wedge_bcs, 0.9, flux_calculator="ausmdv")
_add("naca_omesh", lambda: su2.naca0012_omesh(32, 12),
     lambda: {"wall": cfg.adiabatic_no_slip_wall(), "inflow": cfg.subsonic_inflow(FS2_SUB),
              "outflow": cfg.subsonic_outflow(101325.0)}, 0.2, flux_calculator="ausmdv", viscous=True)



# the two remaining boundary actions: boundary_layer_profile (a cubic spline in
# the ghost centroid's y) and constant_flux (the one FluxAction).  The profile knots keep every ghost
# centroid out of the LAST TWO spline intervals, where the reference reads past the end of its y''
# array (see spline_eval in ibis_b200/csrc/fv_bodies.h): below the first knot, inside, and above the
# last knot are all exercised.
def quad2_profile_bcs():
    h = [0.08, 0.2, 0.3, 0.45, 0.6, 0.8, 1.1, 2.0, 3.0]                 # ghost y in (0, 1): clamp below 0.08
    v = [0.0, 310.0, 520.0, 700.0, 840.0, 930.0, 990.0, 1000.0, 1000.0]
    T = [500.0, 470.0, 430.0, 390.0, 350.0, 320.0, 305.0, 300.0, 300.0]
    b = quad2_visc_bcs()
    b["left"] = cfg.boundary_layer_inflow(h, v, T, 101325.0)
    return b


def box3_profile_flux_bcs():
    # regular 6^3 box: ghost centroids at y = (j + 1/2)/6; 0.5833 is in [0.45, 0.6], none in the last two
    # intervals [0.6, 0.65], [0.65, 0.7]; 0.75 and 0.9167 take the clamp above the last knot
    h = [0.0, 0.1, 0.2, 0.3, 0.45, 0.6, 0.65, 0.7]
    v = [0.0, 260.0, 480.0, 650.0, 820.0, 930.0, 950.0, 960.0]
    T = [450.0, 430.0, 400.0, 370.0, 335.0, 310.0, 304.0, 300.0]
    b = box3_bcs()
    b["west"] = cfg.boundary_layer_inflow(h, v, T, 95000.0)
    b["south"] = cfg.constant_flux(FS3)
    b["bottom"] = cfg.slip_wall()
    return b


def quad2_flux_visc_bcs():
    b = quad2_visc_bcs()
    b["left"] = cfg.constant_flux(FS2)
    b["top"] = cfg.constant_flux(cfg.flow_state(rho=0.9, T=350.0, vx=-300.0, vy=-450.0))
    return b


_add("quad2d_bl_profile", _quad, quad2_profile_bcs, 0.5, flux_calculator="ausmdv", viscous=True)
_add("box3d_profile_cflux", lambda: su2.box_hex(6), box3_profile_flux_bcs, flux_calculator="ausmdv")
_add("quad2d_cflux_viscous", _quad, quad2_flux_visc_bcs, 0.5, flux_calculator="hanel", viscous=True)

# moving-grid residuals (first part): a case of CASES + a rigid-body translation
# of the grid (finite_volume/rigid_body_translation.cpp) -> fixtures tests/golden/moving_<name>.npz
MOVING_CASES = {
    "box3d_ausmdv": (120.0, -75.0, 40.0),
    "box3d_rusanov": (-310.0, 20.0, 55.0),
    "quad2d_hanel": (-200.0, 90.0, 0.0),
    "quad2d_ldfss2": (75.0, 130.0, 0.0),
    "mixed3d_walls": (50.0, 60.0, -70.0),          # triangular faces: (v + v + v) / 3 is not always v
    "tet3d": (33.0, -21.0, 12.0),
    "quad2d_cflux_viscous": (150.0, -30.0, 0.0),   # constant_flux relative to the moving interface, viscous
    "box3d_profile_cflux": (-40.0, 25.0, 10.0),
}


def build_moving_case(name):
    """-> (mesh, config with grid.motion enabled, Grid, prim, vertex velocities (n_vertices, 3))"""
    mesh_f, bcs_f, mach, kw = CASES[name]
    mesh = mesh_f()
    bcs = bcs_f()
    vel = MOVING_CASES[name]
    config = cfg.make_config(mesh.dim, bcs, motion=cfg.rigid_body_translation(*vel), **kw)
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    prim = perturbed_state(g, mesh.dim, mach)
    vvel = np.tile(np.array(vel, dtype=np.float64), (len(mesh.vertices), 1))
    return mesh, config, g, prim, vvel


VISCOUS_CASES = [k for k, v in CASES.items() if v[3].get("viscous")]


def build_case(name):
    """-> (mesh, config dict, Grid, prim (n_total, 7))"""
    mesh_f, bcs_f, mach, kw = CASES[name]
    mesh = mesh_f()
    bcs = bcs_f()
    config = cfg.make_config(mesh.dim, bcs, **kw)
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    prim = perturbed_state(g, mesh.dim, mach)
    return mesh, config, g, prim


def dual_seed(prim, seed=99):
    """A deterministic dual part for a primitive state (JFNK-style perturbation direction)."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(prim * rng.uniform(-1e-3, 1e-3, prim.shape))
