"""The boundary_layer_profile ghost-cell action with its CubicSpline, and the constant_flux flux action
(finite_volume/boundaries/boundary.cpp:51-95, 405-444, util/cubic_spline.*).

* the reference's own known-answer tests of CubicSpline (util/cubic_spline.cpp:91-160: "CubicSpline",
  "CubicSpline2"), driven through the ghost fill: the action evaluates the spline at the y of each
  ghost-cell centroid, and the centroids are an INPUT of the C ABI, so the test places them on the
  reference's query points;
* a numpy restatement of ConstantFlux::apply for a cell that has nothing but constant-flux faces;
* the error behaviour of the reference's two factories (boundary.cpp:449-492).

The full-residual parity of both actions against the compiled reference is in test_emu_parity.py /
test_gpu_parity.py (cases quad2d_bl_profile, box3d_profile_cflux, quad2d_cflux_viscous).
Each test body runs on the host-emulation build here and on the CUDA build under -m gpu."""
import copy

import numpy as np
import pytest

import cases
from emu_util import emu_lib
from ibis_b200 import config as cfg
from ibis_b200 import grid as gridmod
from ibis_b200 import su2
from ibis_b200.lib import Context, IbisB200Error

BUILDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _lib(build):
    if build == "emu":
        return emu_lib()
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return None


def _profile_ghost_fill(build, x, y, queries):
    """ghost (T, vx) of a boundary_layer_profile with v = T = spline(x, y), the ghost centroids of the
    `left` boundary moved to `queries`"""
    mesh = su2.box_quad(4, len(queries))
    bcs = cases.quad2_bcs()
    bcs["left"] = cfg.boundary_layer_inflow(x, y, y, 101325.0)
    config = cfg.make_config(2, bcs, flux_calculator="ausmdv")
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    ghosts = np.asarray(g.tags["left"]["ghosts"])
    assert len(ghosts) == len(queries)
    g.cell_centroid = np.array(g.cell_centroid)
    g.cell_centroid[ghosts, 1] = queries
    prim = cases.perturbed_state(g, 2)
    ctx = Context(g, config, lib_path=_lib(build))
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.compute_dudt(fs, dq, True)
    out = ctx.fs_download(fs)[ghosts]
    ctx.close()
    # update_thermo_from_pT and the zeroed transverse velocity (boundary.cpp:86-93)
    assert np.array_equal(out[:, 1], np.full(len(ghosts), 101325.0))
    with np.errstate(divide="ignore"):
        assert np.array_equal(out[:, 0], 101325.0 / (287.0 * out[:, 2]))
    assert np.array_equal(out[:, 3], 717.5 * out[:, 2])
    assert not out[:, 5:].any()
    assert np.array_equal(out[:, 2], out[:, 4])           # same table for T and v
    return out[:, 4]


@pytest.mark.parametrize("build", BUILDS)
def test_cubic_spline_reference_kat_1(build):
    # util/cubic_spline.cpp:91-121
    got = _profile_ghost_fill(build, [1.0, 2.0, 3.0, 4.0, 8.0], [4.0, 6.0, 8.0, 10.0, 18.0],
                              [1.0, 1.5, 2.0, 4.0, 5.0, 0.0, 7.0])
    np.testing.assert_allclose(got, [4.0, 5.0, 6.0, 10.0, 12.0, 4.0, 16.0], rtol=1e-12)


@pytest.mark.parametrize("build", BUILDS)
def test_cubic_spline_reference_kat_2(build):
    # util/cubic_spline.cpp:123-160
    x = [0.000000000, 0.000286575, 0.000642487, 0.001041700, 0.001466750, 0.001900160,
         0.002326750, 0.003120560, 0.003816400, 0.004433000, 0.005547930, 0.006617530]
    y = [0.00000, 153.017, 319.372, 488.652, 654.953, 811.480, 951.827, 1167.76, 1295.30, 1356.55, 1387.81, 1390.00]
    got = _profile_ghost_fill(build, x, y, [0.0, 0.000286575, 0.000642487, 0.005547930, 0.1, -0.1])
    np.testing.assert_allclose(got, [0.0, 153.017, 319.372, 1387.81, 1390.0, 0.0], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("build", BUILDS)
def test_cubic_spline_is_the_natural_spline(build):
    """interpolates the knots, C2 across them, zero curvature at both ends: checked against scipy's
    natural cubic spline, including the last two intervals (where the reference itself is undefined)"""
    from scipy.interpolate import CubicSpline
    x = np.array([0.0, 0.1, 0.25, 0.3, 0.5, 0.8, 1.0])
    y = np.array([0.0, 300.0, 520.0, 600.0, 810.0, 950.0, 1000.0])
    q = np.array([0.05, 0.1, 0.2, 0.27, 0.31, 0.45, 0.6, 0.79, 0.85, 0.99])
    got = _profile_ghost_fill(build, list(x), list(y), q)
    np.testing.assert_allclose(got, CubicSpline(x, y, bc_type="natural")(q), rtol=1e-12)


@pytest.mark.parametrize("build", BUILDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_constant_flux_restated(build, dim):
    """A one-cell grid whose faces all carry constant_flux: dudt = -sum_f F_f A_f outsign / V with
    F_f written as in ConstantFlux::apply (boundary.cpp:405-444), including its `vy * n.x` in the
    energy flux; independent of the flow state in the cell."""
    fs = cfg.flow_state(rho=0.8, T=410.0, vx=620.0, vy=-340.0, vz=150.0 if dim == 3 else 0.0)
    if dim == 2:
        mesh = su2.box_quad(1, 1, lengths=(0.7, 0.4))
        names = ["left", "right", "bottom", "top"]
    else:
        mesh = su2.box_hex(1, 1, 1, lengths=(0.7, 0.4, 0.9))
        names = ["west", "east", "south", "north", "bottom", "top"]
    bcs = {t: cfg.constant_flux(fs) for t in names}
    config = cfg.make_config(dim, bcs, flux_calculator="hanel")
    g = gridmod.build_grid(mesh, names)
    prim = cases.perturbed_state(g, dim)
    ctx = Context(g, config, lib_path=_lib(build))
    f, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(f, prim)
    n = np.asarray(g.face_norm)
    mf = fs["rho"] * (fs["vx"] * n[:, 0] + fs["vy"] * n[:, 1] + fs["vz"] * n[:, 2])
    p, u = fs["p"], 717.5 * fs["T"]
    F = [mf, p * n[:, 0] + fs["vx"] * mf, p * n[:, 1] + fs["vy"] * mf]
    if dim == 3:
        F.append(p * n[:, 2] + fs["vz"] * mf)
    F.append(mf * (u + 0.5 * (fs["vx"] * fs["vx"] + fs["vy"] * fs["vy"] + fs["vz"] * fs["vz"])) +
             p * (fs["vx"] * n[:, 0] + fs["vy"] * n[:, 0] + fs["vz"] * n[:, 2]))
    F = np.stack(F, axis=1)
    want = np.zeros(dim + 2)
    for j in range(int(g.cell_face_offsets[1])):                 # finite_volume.cpp:221-256, in cell-face order
        fid = int(g.cell_face_ids[j])
        want = want + F[fid] * (-g.face_area[fid] * g.cell_outsigns[j])
    want = want / g.cell_volume[0]
    for allow in (True, False):
        ctx.compute_dudt(f, dq, allow)
        np.testing.assert_array_equal(ctx.cq_download(dq)[0], want)
    ctx.close()


def test_action_factories_reject_like_the_reference():
    """build_boundary_action / build_flux_action (boundary.cpp:449-492): a ghost-cell action in the flux
    list or a flux action in a ghost-cell list is unknown to that factory; CubicSpline needs 4 points and
    lists of one length (cubic_spline.cpp:12-19)."""
    lib = emu_lib()
    mesh = su2.box_quad(3, 3)
    names = ["left", "right", "bottom", "top"]
    g = gridmod.build_grid(mesh, names)

    def make(mutate):
        bcs = cases.quad2_bcs()
        mutate(bcs)
        return Context(g, cfg.make_config(2, bcs), lib_path=lib)

    with pytest.raises(IbisB200Error, match="Unknown flux action"):
        make(lambda b: b["left"]["post_convective_flux"].append({"type": "internal_copy"}))
    with pytest.raises(IbisB200Error, match="Unknown boundary action"):
        make(lambda b: b["left"]["pre_reconstruction"].append({"type": "constant_flux", "flow_state": dict(cases.FS2)}))
    with pytest.raises(IbisB200Error, match="Insufficient points for cubic spline"):
        make(lambda b: b.update(left=cfg.boundary_layer_inflow([0.0, 0.5, 1.0], [0.0, 1.0, 2.0], [3.0, 3.0, 3.0], 1e5)))
    with pytest.raises(IbisB200Error, match="different number of x and y points"):
        make(lambda b: b.update(left=cfg.boundary_layer_inflow([0.0, 0.3, 0.6, 1.0], [0.0, 1.0, 2.0], [3.0] * 4, 1e5)))
    # and the library itself (not only the Python flattening) refuses a misplaced action
    from ibis_b200 import lib as libmod
    bcs = cases.quad2_bcs()
    config = cfg.make_config(2, bcs)
    orig = libmod.flatten_config

    def swapped(config_, tags):
        c = orig(config_, tags)
        c.bc_actions[0].type = libmod.FLUX_ACTION["constant_flux"]      # a flux action in pre_reconstruction
        return c

    libmod.flatten_config = swapped
    try:
        with pytest.raises(IbisB200Error, match="Unknown boundary action"):
            Context(g, copy.deepcopy(config), lib_path=lib)
    finally:
        libmod.flatten_config = orig
