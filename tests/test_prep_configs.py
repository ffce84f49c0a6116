"""job.py -> the reference's own `ibis prep` -> config/config.json -> THIS repository's configuration flattening
(ibis_b200/lib.py::flatten_config) and host solvers (ibis_b200/host: Solver::solve / RungeKutta / SteadyState / Jfnk /
Gmres above the C ABI), with none of the reference's code in the loop: "the configuration interface is unchanged".

The prepared cases are the fixtures of tests/golden/dropin (tests/golden/make_dropin_cases.py: the reference's
tests/minimal cases and three of its examples, job.py unchanged through the reference's prep.py:1178-1251); the final
snapshot must equal the one the reference's own solver wrote: bit for bit for explicit inviscid runs, 1e-11 with
viscous terms (pow in Sutherland's law), 1e-9 for the steady run (the reductions of its Krylov solver are summed in
another order).  Host emulation here, the B200 under -m gpu."""
import json
import os

import numpy as np
import pytest

from ibis_b200 import host, native_io
from ibis_b200 import lib as libmod
from ibis_b200.lib import Context

DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin")
CASES = {"uniform_stream": 0.0, "uniform_stream_3D": 0.0, "steady_state_uniform_stream": 1e-9,
         "example_wedge_transient": 0.0, "example_forward_facing_step": 0.0, "example_naca0012_viscous": 1e-11}


def _run(case, lib_path=None, host_lib=None):
    with open(os.path.join(DROPIN, case, "config", "config.json")) as f:
        config = json.load(f)                              # exactly what the reference's prep wrote (+ the listed edits)
    mesh = libmod.read_su2(os.path.join(DROPIN, case, "io", "grid", "0000", "block_0000.su2"), lib_path=lib_path)
    g = libmod.build_grid(mesh, None, lib_path=lib_path)
    assert sorted(config["grid"]["boundaries"]) == list(g.tags)
    exp = np.load(os.path.join(DROPIN, case, "expected.npz"))
    ctx = Context(g, config, lib_path=lib_path)
    try:
        fs = ctx.fs_create()
        native_io.read_snapshot(ctx, fs, os.path.join(DROPIN, case, "io", "flow"), 0)     # snapshot 0, as prep left it
        prim0 = ctx.fs_download(fs)
        n_plots = []
        out, rep = host.solve(ctx, config, prim0, host_lib=host_lib, on_plot=lambda step, t, st: n_plots.append(step))
    finally:
        ctx.close()
    assert rep["status"] == int(exp["rc"]) == 0
    cols = {"T": 2, "p": 1, "vx": 4, "vy": 5, "vz": 6}
    n, rtol = g.n_cells, CASES[case]
    for k, col in cols.items():
        if "final_" + k not in exp:
            continue
        want = exp["final_" + k]
        if rtol == 0.0:
            assert np.array_equal(out[:n, col], want), f"{case}: {k} is not the reference's snapshot"
        else:
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(out[:n, col] - want).max() <= rtol * scale, f"{case}: {k}"
    if config["solver"]["name"] == "runge_kutta":
        assert abs(rep["t"] - float(exp["final_time"])) <= 1e-12 * float(exp["final_time"])


@pytest.mark.parametrize("case", sorted(CASES))
def test_prepared_reference_cases_on_the_host_emulation(case):
    from emu_util import emu_host_lib, emu_lib
    _run(case, emu_lib(), emu_host_lib())


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_prepared_reference_cases_on_the_gpu(case):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _run(case)


def _naca_steady(lib_path=None, host_lib=None):
    """BASELINE.json configs[2]'s steady variant: examples/naca0012/job.py with SteadyState(...) swapped in for
    RungeKutta(...), prepared by the reference's own prep (viscous, subsonic in/outflow, adiabatic no-slip wall, JFNK +
    GMRES(30)).  The reference itself cannot run it (heap corruption before its first step completes:
    tests/golden/make_dropin_cases.py), so there is no snapshot to compare with; the product must run its twelve
    Newton steps, keep every cell physical and bring the global residual down."""
    case = "example_naca0012_steady"
    with open(os.path.join(DROPIN, case, "config", "config.json")) as f:
        config = json.load(f)
    assert config["solver"]["name"] == "steady_state" and config["viscous_flux"]["enabled"]
    mesh = libmod.read_su2(os.path.join(DROPIN, case, "io", "grid", "0000", "block_0000.su2"), lib_path=lib_path)
    g = libmod.build_grid(mesh, None, lib_path=lib_path)
    ctx = Context(g, config, lib_path=lib_path)
    try:
        fs = ctx.fs_create()
        native_io.read_snapshot(ctx, fs, os.path.join(DROPIN, case, "io", "flow"), 0)
        prim0 = ctx.fs_download(fs)
        norms = []
        out, rep = host.solve(ctx, config, prim0, host_lib=host_lib, on_residuals=lambda step, t, nm: norms.append(nm[0]))
        fs2 = ctx.fs_create()
        ctx.fs_upload(fs2, out)
        assert ctx.count_bad_cells(fs2) == 0
    finally:
        ctx.close()
    assert rep["status"] == 0 and rep["steps"] == config["solver"]["max_steps"]
    n = g.n_cells
    assert np.isfinite(out[:n]).all() and (out[:n, 0] > 0).all() and (out[:n, 2] > 0).all()
    assert len(norms) >= 2 and norms[-1] < 0.5 * norms[0], norms


def test_naca0012_steady_state_runs_on_the_host_emulation():
    from emu_util import emu_host_lib, emu_lib
    _naca_steady(emu_lib(), emu_host_lib())


@pytest.mark.gpu
def test_naca0012_steady_state_runs_on_the_gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _naca_steady()
