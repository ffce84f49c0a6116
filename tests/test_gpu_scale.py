"""GPU tests (-m gpu) at sizes where a thread block no longer sees the whole grid (tens of thousands of
tiles): properties that do not need a stored answer.

* the two independent implementations of the second-order residual -- the sides path (tile-owned
  interface records, TMA) and the gather path (gradient / limiter arrays, per-tile gathers) -- agree
  BIT FOR BIT on a 2 M-cell perturbed box, and so do the dual evaluations' real parts;
* the unmodified reference (oracle/_ref, when it travelled to the box) agrees bit for bit on a 64^3 box
  it can build in seconds;
* encode -> decode: native snapshot in, snapshot out, conversions there and back."""
import os
import sys

import numpy as np
import pytest

import cases
from ibis_b200 import config as cfg
from ibis_b200.boxgrid import BoxGrid
from ibis_b200.lib import Context

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config(flux="ausmdv"):
    return cfg.make_config(3, cases.box3_bcs(), flux_calculator=flux, reconstruction_order=2,
                           solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5))


def _residual(g, prim, config, monkeypatch, sides, dual=False):
    monkeypatch.setenv("IBIS_B200_SIDES", sides)
    ctx = Context(g, config)
    fs, dq = ctx.fs_create(dual=dual), ctx.cq_create(dual=dual)
    ctx.fs_upload(fs, prim, 1e-3 * prim if dual else None)
    ctx.compute_dudt(fs, dq, True)
    out = ctx.cq_download(dq, dual=dual)
    dt = ctx.estimate_dt(fs)
    ctx.close()
    return out, dt


def test_sides_path_equals_gather_path_on_a_box_that_is_not_a_power_of_two(monkeypatch):
    """200^3 = 8 M cells (the per-rank size of the 8-GPU bench): the Morton bricks do not tile the box evenly"""
    import torch
    assert torch.cuda.is_available()
    g = BoxGrid(200)
    prim = cases.perturbed_state(g, 3)
    a, dta = _residual(g, prim, _config("ausmdv"), monkeypatch, "1")
    b, dtb = _residual(g, prim, _config("ausmdv"), monkeypatch, "0")
    assert np.array_equal(a, b) and dta == dtb
    assert np.isfinite(a[:g.n_cells]).all()


@pytest.mark.parametrize("flux", ["ausmdv", "ldfss2"])
def test_sides_path_equals_gather_path_on_two_million_cells(flux, monkeypatch):
    import torch
    assert torch.cuda.is_available()
    g = BoxGrid(128)
    prim = cases.perturbed_state(g, 3)
    a, dta = _residual(g, prim, _config(flux), monkeypatch, "1")
    b, dtb = _residual(g, prim, _config(flux), monkeypatch, "0")
    assert np.array_equal(a, b) and dta == dtb
    assert np.isfinite(a[:g.n_cells]).all() and np.abs(a[:g.n_cells]).max() > 0
    if flux == "ausmdv":
        (re_, du_), _ = _residual(g, prim, _config(flux), monkeypatch, "1", dual=True)
        assert np.array_equal(re_[:g.n_cells], a[:g.n_cells]), "real part of the dual evaluation"
        assert np.isfinite(du_[:g.n_cells]).all()


@pytest.mark.parametrize("n", [64, 128])
def test_reference_itself_on_a_whole_box(n, monkeypatch):
    """The compiled reference (Kokkos-OpenMP, oracle/_ref) evaluates the bench configuration on the n^3 box on this
    machine's host cores; the CUDA residual must equal it bit for bit.  128^3 = 2.1 M cells is the box the bench's
    reference arm times."""
    import torch
    assert torch.cuda.is_available()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refharness as rh
    if not rh.available():
        pytest.skip("oracle/_ref/libibis_ref.so did not travel to this box")
    from ibis_b200 import grid as gridmod
    from ibis_b200 import su2
    mesh = su2.box_hex(n)
    g = gridmod.build_grid(mesh, ["west", "east", "south", "north", "bottom", "top"])
    path = "/tmp/test_scale_box.su2"
    mesh.write(path)
    prim = cases.perturbed_state(g, 3)
    rc = rh.RefCase(path, _config())
    want = rc.compute_dudt(prim)
    rc.close()
    want = want[0] if isinstance(want, tuple) else want
    got, _ = _residual(g, prim, _config(), monkeypatch, "1")
    assert np.array_equal(got[:g.n_cells], np.asarray(want)[:g.n_cells])


def test_snapshot_and_conversion_round_trips(monkeypatch):
    import torch
    assert torch.cuda.is_available()
    g = BoxGrid(96)
    prim = cases.perturbed_state(g, 3)
    n = g.n_cells
    ctx = Context(g, _config())
    fs, cq, fs2 = ctx.fs_create(), ctx.cq_create(), ctx.fs_create()
    T, p = np.ascontiguousarray(prim[:n, 2]), np.ascontiguousarray(prim[:n, 1])
    v = [np.ascontiguousarray(prim[:n, 4 + k]) for k in range(3)]
    ctx.flow_read_native(fs, T, p, *v)
    back = ctx.flow_write_native(fs)
    for a, b in zip(back, [T, p] + v):
        assert np.array_equal(a, b)
    ctx.prim_to_cons(cq, fs)
    ctx.cons_to_prim(cq, fs2)
    a, b = ctx.fs_download(fs)[:n], ctx.fs_download(fs2)[:n]
    np.testing.assert_allclose(b, a, rtol=4e-15)          # rho*v/rho etc.: a few ulp, not exact
    ctx.close()
