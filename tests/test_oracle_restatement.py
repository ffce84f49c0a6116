"""oracle/restatement.py (plain numpy, written from the reference's sources) against the golden vectors of the
COMPILED reference: bit for bit -- residual, ghost fill, every gradient and limiter, second and first order --
on every inviscid case it covers.  This pins the restatement; tests/test_gpu_parity.py then holds the CUDA path
to the same vectors, and test_product_equals_restatement_on_seeded_input below holds it to the restatement on
inputs no fixture stores."""
import os
import sys

import numpy as np
import pytest

import cases
import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import restatement  # noqa: E402

COVERED = sorted(cases.CASES)             # every case of tests/cases.py


def test_coverage_is_what_the_header_says():
    assert len(COVERED) == 26 and {"box3d_ausmdv", "box3d_hanel", "box3d_rusanov", "box3d_ldfss2", "quad2d_ldfss2",
                                    "quad2d_subsonic_bcs", "box3d_profile_cflux", "quad2d_ausmdv", "mixed3d_walls",
                                    "tet3d", "tri2d", "wedge2d", "box3d_rho_p", "box3d_rho_T", "box3d_unlimited",
                                    "box3d_order1"} <= set(COVERED)


@pytest.mark.parametrize("name", COVERED)
def test_restatement_equals_compiled_reference(name):
    mesh, config, g, prim = cases.build_case(name)
    gold = parity.load_golden(name)
    for allow, sfx in ((True, ""), (False, "_o1")):
        dudt, fs, inter = restatement.compute_dudt(g, config, prim, allow)
        inter.pop("face_fs", None)
        parity.assert_close(dudt, gold["dudt" + sfx], 1.0, f"{name}: dudt{sfx}", True)
        parity.assert_close(fs, gold["prim_out" + sfx], 1.0, f"{name}: ghost fill{sfx}", True)
        if allow:
            for k, v in inter.items():
                if k in gold:
                    parity.assert_close(v, gold[k], 1.0, f"{name}: {k}", True)
            if config["convective_flux"]["reconstruction_order"] == 2:
                assert any(k.startswith("grad_") for k in inter)


@pytest.mark.parametrize("name", sorted(cases.MOVING_CASES))
def test_restatement_moving_grid_equals_compiled_reference(name):
    """the moving-grid overload (rigid_body_translation driver) against tests/golden/moving_*.npz"""
    mesh, config, g, prim, vvel = cases.build_moving_case(name)
    gold = dict(np.load(os.path.join(parity.GOLDEN, "moving_" + name + ".npz")))
    cq = restatement.primatives_to_conserved(prim, mesh.dim)
    for allow, sfx in ((True, ""), (False, "_o1")):
        dudt, fs, inter = restatement.compute_dudt(g, config, prim, allow, vertex_vel=vvel, cq=cq)
        parity.assert_close(dudt, gold["dudt" + sfx], 1.0, f"{name}: moving dudt{sfx}", True)
        parity.assert_close(fs, gold["prim_out" + sfx], 1.0, f"{name}: ghost fill{sfx}", True)
        parity.assert_close(inter["face_vel"], gold["face_vel" + sfx], 1.0, f"{name}: face_vel{sfx}", True)


@pytest.mark.parametrize("name", ["rk_wedge2d", "rk_box3d", "rk_quad2d_cflux", "rk_ffs2d", "rk_plate2d_viscous",
                                  "rk_quad2d_moving", "rk_box3d_moving"])
def test_restatement_runge_kutta_equals_reference_snapshot(name):
    """the explicit solver restated (Solver::solve + RungeKutta, aliased stage arrays included) against the final
    snapshot of the reference's own run (tests/golden/solve_*.npz): bit for bit"""
    import solve_cases as sc
    mesh, config, g, prim0, snap = sc.build_solve_case(name)
    gold = dict(np.load(os.path.join(parity.GOLDEN, "solve_" + name + ".npz")))
    flow, t, steps = restatement.runge_kutta(g, config, prim0, mesh)
    n = g.n_cells
    assert t == float(gold["final_time"])
    for k, col in (("T", 2), ("p", 1), ("vx", 4), ("vy", 5), ("vz", 6)):
        if "final_" + k in gold:
            assert np.array_equal(flow[:n, col], gold["final_" + k]), f"{name}: {k} after {steps} steps"


def test_restatement_cubic_spline_reference_kats():
    """util/cubic_spline.cpp:91-160"""
    s = restatement.CubicSpline([1.0, 2.0, 3.0, 4.0, 8.0], [4.0, 6.0, 8.0, 10.0, 18.0])
    np.testing.assert_allclose([s.eval(q) for q in (1.0, 1.5, 2.0, 4.0, 5.0, 0.0, 7.0)], [4, 5, 6, 10, 12, 4, 16], rtol=1e-12)
    x = [0.000000000, 0.000286575, 0.000642487, 0.001041700, 0.001466750, 0.001900160,
         0.002326750, 0.003120560, 0.003816400, 0.004433000, 0.005547930, 0.006617530]
    y = [0.00000, 153.017, 319.372, 488.652, 654.953, 811.480, 951.827, 1167.76, 1295.30, 1356.55, 1387.81, 1390.00]
    s = restatement.CubicSpline(x, y)
    np.testing.assert_allclose([s.eval(q) for q in (0.0, 0.000286575, 0.000642487, 0.005547930, 0.1, -0.1)],
                               [0.0, 153.017, 319.372, 1387.81, 1390.0, 0.0], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", COVERED)
def test_restatement_scalars_and_conversions(name):
    """estimate_dt, count_bad_cells, L2 norms, primitive <-> conserved against the compiled reference's vectors"""
    mesh, config, g, prim = cases.build_case(name)
    gold = parity.load_golden(name)
    face_fs = restatement.compute_dudt(g, config, prim, True)[2].get("face_fs")
    assert restatement.estimate_dt(g, config, gold["prim_out"], face_fs) == float(gold["dt"])
    assert restatement.count_bad_cells(g, gold["prim_out"]) == int(gold["bad"])
    cons = restatement.primatives_to_conserved(prim, mesh.dim)
    parity.assert_close(cons, gold["cons"], 1.0, f"{name}: prim_to_cons", True)
    back = restatement.conserved_to_primatives(gold["cons"], mesh.dim, config["gas_model"])
    parity.assert_close(back, gold["prim_from_cons"], 1.0, f"{name}: cons_to_prim", True)
    # the reference sums over all rows of dudt (ghost rows are zero); the fixture was made with its default
    # thread count, so the sums agree to round-off, not bit for bit
    np.testing.assert_allclose(restatement.l2_norms(gold["dudt"], mesh.dim), gold["l2"], rtol=1e-13)


@pytest.mark.parametrize("build", [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("seed", [7, 20251018])
def test_product_equals_restatement_on_seeded_input(build, seed):
    """a larger jittered box and a state no fixture holds: the product path through the C ABI (host harness here,
    CUDA under -m gpu) == restatement, bit for bit"""
    lib_path = None
    if build == "emu":
        from emu_util import emu_lib
        lib_path = emu_lib()
    else:
        import torch
        assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ibis_b200 import config as cfg
    from ibis_b200 import grid as gridmod
    from ibis_b200 import su2
    from ibis_b200.lib import Context
    mesh = su2.box_hex(14, 12, 10, jitter=0.25, seed=seed)
    bcs = cases.box3_wall_bcs()
    config = cfg.make_config(3, bcs, flux_calculator="ausmdv")
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    prim = cases.perturbed_state(g, 3, 0.7, seed=seed)
    ctx = Context(g, config, lib_path=lib_path)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    for allow in (True, False):
        ctx.fs_upload(fs, prim)
        ctx.compute_dudt(fs, dq, allow)
        want, ghosts, _ = restatement.compute_dudt(g, config, prim, allow)
        assert np.array_equal(ctx.cq_download(dq)[:g.n_cells], want)
        assert np.array_equal(ctx.fs_download(fs), ghosts)
    ctx.close()


@pytest.mark.parametrize("name,allow", [("box3d_unlimited", True), ("box3d_hanel", False), ("quad2d_rusanov", False)])
def test_dual_residual_is_the_derivative_of_the_restated_residual(name, allow):
    """ties the dual-number path (the exact Jacobian-vector product of the steady solver) to the restatement, which
    has no dual numbers: dual part of R(u + eps v) == central difference of the restated real residual along v.
    Smooth configurations only (unlimited second order, first order): a limiter is not differentiable."""
    from emu_util import emu_lib
    from ibis_b200.lib import Context
    mesh, config, g, prim = cases.build_case(name)
    v = cases.dual_seed(prim)
    ctx = Context(g, config, lib_path=emu_lib())
    fsd, dqd = ctx.fs_create(True), ctx.cq_create(True)
    ctx.fs_upload(fsd, prim, v)
    ctx.compute_dudt(fsd, dqd, allow)
    _, du = ctx.cq_download(dqd, dual=True)
    ctx.close()
    h = 1e-4                                     # v is ~1e-3 of the state: the step is 1e-7 relative
    plus = restatement.compute_dudt(g, config, prim + h * v, allow)[0]
    minus = restatement.compute_dudt(g, config, prim - h * v, allow)[0]
    fd = (plus - minus) / (2.0 * h)
    scale = np.abs(du[:g.n_cells]).max(axis=0)
    assert (np.abs(du[:g.n_cells] - fd) <= 2e-5 * scale).all(), np.abs(du[:g.n_cells] - fd).max(axis=0) / scale


@pytest.mark.parametrize("name", COVERED)
def test_restatement_on_dual_numbers_equals_compiled_reference(name):
    """FiniteVolume<dual>::compute_dudt restated (oracle/dualnum.py carries the reference's Dual operators) against
    the compiled reference's dual residual: real AND dual part bit for bit on the inviscid cases (viscous: 1e-12,
    numpy's pow(x, 0.5) in d/dx pow(x, 1.5) is sqrt, the reference's is glibc pow)"""
    from dualnum import Dual
    mesh, config, g, prim = cases.build_case(name)
    gold = parity.load_golden(name)
    exact = not config["viscous_flux"]["enabled"]
    S = parity.flux_scale(g, prim, mesh.dim)
    for allow, sfx in ((True, ""), (False, "_o1")):
        d, fs, _ = restatement.compute_dudt(g, config, Dual(prim, cases.dual_seed(prim)), allow)
        assert isinstance(d, Dual)
        parity.assert_close(d.re, gold["dual_dudt_re" + sfx], S, f"{name}: dual re{sfx}", exact)
        parity.assert_close(d.du, gold["dual_dudt_du" + sfx], 1e-3 * S, f"{name}: dual du{sfx}", exact)
        parity.assert_close(fs.du, gold["dual_prim_out_du" + sfx], 1.0, f"{name}: ghost fill du{sfx}", exact)


@pytest.mark.parametrize("name", sorted(cases.MOVING_CASES))
def test_restatement_moving_grid_on_dual_numbers(name):
    from dualnum import Dual
    mesh, config, g, prim, vvel = cases.build_moving_case(name)
    gold = dict(np.load(os.path.join(parity.GOLDEN, "moving_" + name + ".npz")))
    exact = not config["viscous_flux"]["enabled"]
    S = parity.flux_scale(g, prim, mesh.dim)
    dprim = Dual(prim, cases.dual_seed(prim))
    cq = restatement.primatives_to_conserved(dprim, mesh.dim)
    for allow, sfx in ((True, ""), (False, "_o1")):
        d, _, _ = restatement.compute_dudt(g, config, dprim, allow, vertex_vel=vvel, cq=cq)
        parity.assert_close(d.re, gold["dual_dudt_re" + sfx], S, f"{name}: moving dual re{sfx}", exact)
        parity.assert_close(d.du, gold["dual_dudt_du" + sfx], 1e-3 * S, f"{name}: moving dual du{sfx}", exact)


@pytest.mark.parametrize("name", ["steady_quad2d", "steady_quad2d_cflux", "steady_box3d", "steady_wedge2d",
                                  "steady_plate_fgmres"])
def test_restatement_steady_solver_equals_reference_snapshot(name):
    """JFNK + GMRES restated (dual-number matrix-vector products, Arnoldi, the reference's Givens bookkeeping)
    against the final snapshot of the reference's own steady run, north-star tolerance 1e-9"""
    import solve_cases as sc
    mesh, config, g, prim0, snap = sc.build_solve_case(name)
    gold = dict(np.load(os.path.join(parity.GOLDEN, "solve_" + name + ".npz")))
    flow, steps, rel = restatement.steady_state(g, config, prim0)
    n = g.n_cells
    assert steps - 1 == int(float(gold["final_time"]))
    vscale = np.sqrt(flow[:n, 4] ** 2 + flow[:n, 5] ** 2 + flow[:n, 6] ** 2).max()
    for k, col in (("T", 2), ("p", 1), ("vx", 4), ("vy", 5), ("vz", 6)):
        if "final_" + k in gold:
            b = gold["final_" + k]
            scale = vscale if k.startswith("v") else np.abs(b).max()
            assert np.abs(flow[:n, col] - b).max() <= 1e-9 * scale, (name, k, np.abs(flow[:n, col] - b).max() / scale)


_SWEEP = [
    # mesh factory, dim, bcs factory, make_config kwargs, mach scale
    (lambda s: su2_mod().box_hex(7, 6, 5, jitter=0.3, seed=s), 3, cases.box3_wall_bcs, dict(flux_calculator="hanel", limiter="unlimited"), 0.4),
    (lambda s: su2_mod().box_hex(6, 7, 5, jitter=0.3, seed=s), 3, cases.box3_bcs, dict(flux_calculator="ldfss2", thermo_interpolator="rho_p"), 1.0),
    (lambda s: su2_mod().box_hex(5, 5, 6, jitter=0.2, seed=s), 3, cases.box3_visc_bcs, dict(flux_calculator="rusanov", viscous=True, thermo_interpolator="rho_T"), 0.5),
    (lambda s: su2_mod().mixed_3d(5), 3, cases.box3_wall_bcs, dict(flux_calculator="ausmdv", thermo_interpolator="rho_T"), 0.6),
    (lambda s: su2_mod().box_tet(5), 3, cases.box3_bcs, dict(flux_calculator="rusanov"), 1.0),
    (lambda s: su2_mod().box_quad(14, 11, jitter=0.3, seed=s, triangulate=True), 2, cases.quad2_sub_bcs, dict(flux_calculator="ausmdv"), 0.2),
    (lambda s: su2_mod().box_quad(13, 9, jitter=0.3, seed=s), 2, cases.quad2_visc_bcs, dict(flux_calculator="ldfss2", viscous=True), 0.5),
    (lambda s: su2_mod().box_quad(11, 12, jitter=0.25, seed=s), 2, cases.quad2_flux_visc_bcs, dict(flux_calculator="ausmdv", viscous=True, limiter="unlimited"), 0.5),
]


def su2_mod():
    from ibis_b200 import su2
    return su2


@pytest.mark.parametrize("combo", range(len(_SWEEP)))
@pytest.mark.parametrize("seed", [3, 11])
def test_host_harness_equals_restatement_sweep(combo, seed):
    """configurations and inputs no fixture stores (seeded jitter, seeded state; flux calculator x interpolator x
    limiter x boundary mixes, viscous ones included): the product's kernel bodies through the C ABI (host harness)
    against the restatement -- real residual, ghost fill, and the dual residual's real and dual parts; bit for bit
    where inviscid, 1e-12 where Sutherland's pow enters"""
    from dualnum import Dual
    from emu_util import emu_lib
    from ibis_b200 import config as cfg
    from ibis_b200 import grid as gridmod
    from ibis_b200.lib import Context
    mesh_f, dim, bcs_f, kw, mach = _SWEEP[combo]
    mesh = mesh_f(seed)
    bcs = bcs_f()
    config = cfg.make_config(dim, bcs, **kw)
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    prim = cases.perturbed_state(g, dim, mach, seed=seed)
    pdu = cases.dual_seed(prim, seed=seed + 1)
    exact = not kw.get("viscous")
    S = parity.flux_scale(g, prim, dim)
    ctx = Context(g, config, lib_path=emu_lib())
    fs, dq, fsd, dqd = ctx.fs_create(), ctx.cq_create(), ctx.fs_create(True), ctx.cq_create(True)
    for allow in (True, False):
        ctx.fs_upload(fs, prim)
        ctx.compute_dudt(fs, dq, allow)
        want, ghosts, _ = restatement.compute_dudt(g, config, prim, allow)
        parity.assert_close(ctx.cq_download(dq)[:g.n_cells], want, S, f"combo {combo}: dudt (allow={allow})", exact)
        parity.assert_close(ctx.fs_download(fs), ghosts, 1.0, f"combo {combo}: ghost fill", True)
        ctx.fs_upload(fsd, prim, pdu)
        ctx.compute_dudt(fsd, dqd, allow)
        dr, dd = ctx.cq_download(dqd, dual=True)
        wd = restatement.compute_dudt(g, config, Dual(prim, pdu), allow)[0]
        parity.assert_close(dr[:g.n_cells], wd.re, S, f"combo {combo}: dual re (allow={allow})", exact)
        parity.assert_close(dd[:g.n_cells], wd.du, 1e-3 * S, f"combo {combo}: dual du (allow={allow})",
                            exact and kw["flux_calculator"] != "ldfss2")
    ctx.close()


_RK_SWEEP = [
    (lambda: su2_mod().box_tet(4), 3, cases.box3_wall_bcs, dict(flux_calculator="rusanov"), "rk4", 0.35, 8, None),
    # mixed_3d has inverted cells: their NEGATIVE local time step wins the reference's min, time runs backwards and
    # the run dies of bad cells after six steps -- in the host solver and in the restatement alike
    (lambda: su2_mod().mixed_3d(4), 3, cases.box3_wall_bcs, dict(flux_calculator="rusanov"), "rk4", 0.35, 8, None),
    (lambda: su2_mod().box_quad(9, 8, jitter=0.2, seed=5, triangulate=True), 2, cases.quad2_bcs,
     dict(flux_calculator="ldfss2", thermo_interpolator="rho_p"), "rk3", 0.3, 10, None),
    (lambda: su2_mod().box_hex(4, 5, 4, jitter=0.2, seed=9), 3, cases.box3_bcs, dict(flux_calculator="ausmdv"), "ssp-rk3", 0.45, 9,
     (60.0, -35.0, 20.0)),
    (lambda: su2_mod().box_quad(8, 7, jitter=0.2, seed=2), 2, cases.quad2_bcs, dict(flux_calculator="hanel", limiter="unlimited"),
     "rk4", 0.3, 7, (-90.0, 45.0, 0.0)),
]


@pytest.mark.parametrize("combo", range(len(_RK_SWEEP)))
def test_host_solver_equals_restated_solver_sweep(combo):
    """the C++ host solver above the C ABI (host harness) against the restated explicit solver on runs no fixture
    stores -- other tableaux (rk3, rk4: the aliased stage arrays matter there), other cells, moving grids: the
    final fields bit for bit"""
    import solve_parity  # noqa: F401  (path setup)
    from emu_util import emu_host_lib, emu_lib
    from ibis_b200 import config as cfg
    from ibis_b200 import grid as gridmod
    from ibis_b200 import host
    from ibis_b200.lib import Context
    import solve_cases as sc
    mesh_f, dim, bcs_f, kw, method, cfl, steps, vel = _RK_SWEEP[combo]
    mesh = mesh_f()
    bcs = bcs_f()
    motion = cfg.rigid_body_translation(*vel) if vel else None
    config = cfg.make_config(dim, bcs, motion=motion, solver=cfg.runge_kutta(method=method, cfl=cfl, max_step=steps, max_time=1.0), **kw)
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    full = cases.perturbed_state(g, dim, 1.0, seed=combo + 40)
    n = g.n_cells
    prim0 = sc.snapshot_state(g, full[:n, 2], full[:n, 1], full[:n, 4], full[:n, 5], full[:n, 6])
    ctx = Context(g, config, lib_path=emu_lib())
    if vel:
        ctx.set_face_vertices(g.face_vertices, len(mesh.vertices))
        ctx.set_cell_vertices(mesh.cell_types, mesh.cell_vertices)
        ctx.set_vertex_positions(mesh.vertices)
        ctx.set_vertex_velocities(np.tile(vel, (len(mesh.vertices), 1)))
    try:
        out, rep = host.solve(ctx, config, prim0, fused_step=False, host_lib=emu_host_lib())
    finally:
        ctx.close()
    want, t, nsteps = restatement.runge_kutta(g, config, prim0, mesh)
    assert rep["steps"] == nsteps and rep["t"] == t
    assert nsteps == steps or (combo == 1 and nsteps == 6 and rep["status"] == 1 and t < 0.0)
    a, b = out[:n], want[:n]
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()


@pytest.mark.parametrize("combo", [0, 1])
def test_host_steady_solver_equals_restated_steady_solver(combo):
    """JFNK / GMRES of the C++ host solver (host harness, dot products in index order) against the restated steady
    solver on runs no fixture stores: 1e-9 of the field scale"""
    import subprocess
    import sys
    code = r'''
import sys
sys.path[:0] = [%r, %r, %r]
import numpy as np
import cases, restatement, solve_cases as sc
from emu_util import emu_host_lib, emu_lib
from ibis_b200 import config as cfg, grid as gridmod, host, su2
from ibis_b200.lib import Context
combo = %d
if combo == 0:
    mesh, bcs, dim = su2.box_quad(9, 7, jitter=0.2, seed=4), cases.quad2_bcs(), 2
    kw = dict(flux_calculator="ausmdv", solver=cfg.steady_state(cfl=4.0, max_steps=6, tolerance=1e-12,
              linear_solver={"type": "gmres", "max_iters": 35, "tol": 1e-3}))
else:
    mesh, bcs, dim = su2.box_hex(4, 3, 4, jitter=0.15, seed=6), cases.box3_wall_bcs(), 3
    kw = dict(flux_calculator="rusanov", thermo_interpolator="rho_p",
              solver=cfg.steady_state(cfl={"type": "residual_based", "growth_threshold": 0.9, "power": 1.0, "start_cfl": 1.0,
                                           "max_cfl": 50.0}, max_steps=5, tolerance=1e-12,
                                      linear_solver={"type": "gmres", "max_iters": 30, "tol": 1e-3}))
config = cfg.make_config(dim, bcs, **kw)
g = gridmod.build_grid(mesh, list(bcs.keys()))
full = cases.perturbed_state(g, dim, 1.0, seed=77 + combo)
n = g.n_cells
prim0 = sc.snapshot_state(g, full[:n, 2], full[:n, 1], full[:n, 4], full[:n, 5], full[:n, 6])
ctx = Context(g, config, lib_path=emu_lib())
out, rep = host.solve(ctx, config, prim0, host_lib=emu_host_lib())
ctx.close()
want, steps, rel = restatement.steady_state(g, config, prim0)
assert rep["steps"] == steps, (rep["steps"], steps)
vs = np.sqrt((want[:n, 4:7] ** 2).sum(axis=1)).max()
for col in (1, 2, 4, 5, 6):
    scale = vs if col >= 4 else np.abs(want[:n, col]).max()
    err = np.abs(out[:n, col] - want[:n, col]).max() / scale
    assert err <= 1e-9, (col, err)
print("STEADY_SWEEP_OK", steps)
''' % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), combo)
    env = dict(os.environ, IBIS_B200_ORDERED_DOT="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "STEADY_SWEEP_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
