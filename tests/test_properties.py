"""Size-independent properties of the residual (the checks that still apply where no fixture can be stored),
on the host harness at a size the CPU finishes in seconds:

* conservation: interior fluxes telescope, so in a box closed by slip walls  sum_c V_c dU_c/dt  is zero for mass
  and energy -- to round-off relative to the sum of the magnitudes.  First-order evaluation: there the ghost state
  is the exact mirror image of the wall cell and every flux calculator passes no mass and does no work; with
  reconstruction the ghost carries no gradient, the two sides are no longer mirror images and the reference's
  ghost-cell wall leaks at truncation-error level (a property of the scheme, reproduced, not a defect here);
* free-stream preservation: a uniform supersonic state with matching inflow / extrapolation outflow has zero
  residual through gradients, limiter and all four flux calculators -- on grids whose cells close geometrically
  (regular hexes; jittered quads, whose edges are straight).  Jittered hexes have non-planar faces whose
  two-triangle area and three-vertex normal do not close a cell, and the reference's scheme -- reproduced bit for
  bit -- does not preserve a free stream there;
* the first-order residual does not depend on the limiter or the thermo interpolator."""
import numpy as np
import pytest

import cases
from emu_util import emu_lib
from ibis_b200 import config as cfg
from ibis_b200 import grid as gridmod
from ibis_b200 import su2
from ibis_b200.lib import Context


def _residual(g, config, prim, allow=True):
    ctx = Context(g, config, lib_path=emu_lib())
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.compute_dudt(fs, dq, allow)
    d = ctx.cq_download(dq)[:g.n_cells]
    ctx.close()
    return d


@pytest.mark.parametrize("mesh_f,flux", [(lambda: su2.box_hex(16, 14, 12, jitter=0.25), "ausmdv"),
                                         (lambda: su2.mixed_3d(8), "hanel"), (lambda: su2.box_tet(6), "rusanov")])
def test_closed_box_conserves_mass_and_energy(mesh_f, flux):
    mesh = mesh_f()
    tags = sorted(mesh.markers)
    g = gridmod.build_grid(mesh, tags)
    config = cfg.make_config(3, {t: cfg.slip_wall() for t in tags}, flux_calculator=flux)
    prim = cases.perturbed_state(g, 3, 0.3)
    d = _residual(g, config, prim, False)
    for col in (0, 4):
        contrib = g.cell_volume * d[:, col]
        assert abs(contrib.sum()) <= 1e-11 * np.abs(contrib).sum(), (flux, col)
    d2 = _residual(g, config, prim, True)               # second order: same interior telescoping, leaky walls
    leak = abs((g.cell_volume * d2[:, 0]).sum()) / np.abs(g.cell_volume * d2[:, 0]).sum()
    assert 1e-11 < leak < 0.1


@pytest.mark.parametrize("flux", ["ausmdv", "hanel", "ldfss2", "rusanov"])
@pytest.mark.parametrize("dim", [2, 3])
def test_free_stream_is_preserved(flux, dim):
    if dim == 3:
        mesh, bcs, fs, h = su2.box_hex(12, 10, 8), cases.box3_bcs(), cases.FS3, 1.0 / 12
    else:
        mesh, bcs, fs, h = su2.box_quad(16, 12, jitter=0.25), cases.quad2_bcs(), cases.FS2, 1.0 / 16
        bcs["bottom"] = cfg.supersonic_inflow(fs)              # the flow crosses the bottom: inflow, not a wall
    g = gridmod.build_grid(mesh, sorted(mesh.markers))
    config = cfg.make_config(dim, bcs, flux_calculator=flux)
    d = _residual(g, config, cases.uniform_state(g, dim, fs))
    speed = np.sqrt(fs["vx"] ** 2 + fs["vy"] ** 2 + fs["vz"] ** 2)
    m = fs["rho"] * speed
    scale = np.array([m] + [m * speed] * dim + [m * speed ** 2]) / h
    assert (np.abs(d) <= 1e-11 * scale).all(), np.abs(d).max(axis=0) / scale


def test_first_order_ignores_limiter_and_interpolator():
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    base = _residual(g, config, prim, False)
    for kw in (dict(limiter="unlimited"), dict(thermo_interpolator="rho_p"), dict(thermo_interpolator="rho_T"),
               dict(reconstruction_order=1)):
        other = cfg.make_config(3, cases.box3_bcs(), flux_calculator="ausmdv", **kw)
        assert np.array_equal(_residual(g, other, prim, False), base), kw


def test_an_argument_error_does_not_poison_the_context():
    """A bad handle is reported by the call that passed it (EHANDLE + message) and by nothing after it: the next
    valid call on the same context succeeds (last_error used to stay set and turn every later
    call into ECUDA)."""
    import numpy as np
    import cases
    from emu_util import emu_lib
    from ibis_b200.lib import Context, IbisB200Error
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    ctx = Context(g, config, lib_path=emu_lib())
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    with pytest.raises(IbisB200Error, match="bad handle"):
        ctx.compute_dudt(12345, dq, True)
    with pytest.raises(IbisB200Error):
        ctx.rk_residual_norms()                       # "rk_residual_norms before rk_step"
    ctx.compute_dudt(fs, dq, True)                    # would raise with a poisoned context
    ctx.sync()
    assert np.isfinite(ctx.cq_download(dq)[:g.n_cells]).all()
    assert ctx.count_bad_cells(fs) == 0
    ctx.close()


def test_asynchronous_steps_equal_synchronous_steps():
    """ibis_b200_rk_step with dt == NULL (nothing read back, dt_prev chained on the device) takes the same steps as
    the form that returns dt every step: same dt sequence end point, same state bit for bit."""
    import sys
    import numpy as np
    import cases
    from emu_util import emu_lib
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    tab = cfg.BUTCHER["ssp-rk3"]
    out = []
    for mode in ("sync", "async"):
        ctx = Context(g, config, lib_path=emu_lib())
        fs, cq = ctx.fs_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        dt = sys.float_info.max
        for _ in range(4):
            if mode == "sync":
                dt, dts = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
            else:
                ctx.rk_step_async(fs, cq, tab["a"], tab["b"], 0.5, 1.0)
        if mode == "async":
            dt, dts = ctx.rk_last_dt()
        out.append((dt, dts, ctx.fs_download(fs)))
        ctx.close()
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2])
